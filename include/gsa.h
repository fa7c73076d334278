/*
 * gsa.h -- C ABI of the B200-native suffix-array engine (libgsa.so).
 *
 * This is the drop-in boundary for ONE hot path of fasterthanlime/stringsearch:
 *
 *     divsufsort::sort  ->  sacabase search  ->  sacapart::PartitionedSuffixArray
 *
 * Every entry point cites the reference interface it replaces (file:line relative
 * to the reference checkout).  Signatures use plain pointers and sizes only; no
 * torch / C++ types.  All functions are re-entrant and may be called from several
 * host threads at once (sacapart calls its builder from rayon workers,
 * crates/sacapart/src/lib.rs:41,45-49): every call selects its device and uses its
 * own stream and workspace; the only shared state is a mutex-protected per-device
 * cache of one idle scratch block (see gsa_release_cached_memory).
 *
 * Error convention: the library never aborts.  Codes follow libdivsufsort
 * (c-sources/divsufsort.c:346,359) and extend them:
 *      0  success
 *     -1  invalid arguments (NULL pointer, negative size, ...)
 *     -2  allocation failure (host or device)
 *     -3  CUDA runtime error (gsa_last_error() has the text)
 *     -4  the reference would panic here (empty suffix array, zero partitions...)
 * The language shim on top (Rust / C++ / Python) turns non-zero codes into the
 * panic / exception the reference raises at the same place.
 */
#ifndef GSA_H
#define GSA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSA_OK 0
#define GSA_EINVAL (-1)
#define GSA_ENOMEM (-2)
#define GSA_ECUDA (-3)
#define GSA_EPANIC (-4)

#define GSA_MAX_ROUNDS 40

/* Per-round log of one SA construction (one entry per prefix-doubling round).
 * The harness uses it to recompute the roofline:
 * algorithmic bytes of round k = live * 52 + sorted * 24 * passes + bag * 32
 * (round 0: n * (41 + 24 * passes)). */
typedef struct gsa_round_stat {
  uint64_t depth;      /* symbols of every suffix known to be sorted AFTER this round */
  uint64_t live;       /* live (not yet unique) suffixes walked in this round */
  uint32_t groups;     /* huge groups handled through the group tables in this round */
  uint32_t key_bits;   /* significant bits of the sort key */
  uint32_t passes;     /* 8-bit radix passes actually executed */
  uint32_t sorted;     /* suffixes actually sorted (live minus the inert members of huge groups) */
  float ms_total;      /* device time of the whole round */
  float ms_sort;       /* ... of which radix passes */
  uint32_t bag;        /* suffixes of tiny groups refined outside the sort (the bag) in this round */
  uint32_t reserved_;
} gsa_round_stat;

typedef struct gsa_build_stats {
  uint32_t rounds;
  uint32_t sigma;           /* distinct bytes in the text */
  uint32_t bits_per_symbol; /* code width used for round-0 key packing */
  uint32_t symbols_per_key; /* round-0 depth */
  float ms_total;           /* device time, text resident -> SA resident */
  float ms_h2d, ms_d2h;     /* host<->device copies (host-pointer entry points only) */
  uint64_t radix_pass_launches;
  uint64_t radix_pass_elements; /* sum over pass launches of elements moved */
  float ms_radix_passes;        /* sum of pass kernel time */
  uint64_t kernel_launches;     /* all kernels launched by this build */
  uint64_t radix_pass_bytes;    /* sum over pass launches of bytes read + written: 12 + 12 per element; the first
                                   round-0 pass generates its keys and reads b/8 bytes of packed text per element instead */
  gsa_round_stat round[GSA_MAX_ROUNDS];
} gsa_build_stats;

/* ----------------------------------------------------------------------------
 * SA construction.
 * Replaces:  extern "C" fn divsufsort(T:*const u8, SA:*mut i32, n:i32) -> i32
 *            (crates/cdivsufsort/src/lib.rs:1-3; c-sources/divsufsort.c:331-370,
 *            header c-sources/divsufsort.h) and therefore
 *            divsufsort::sort_in_place / cdivsufsort::sort_in_place
 *            (crates/divsufsort/src/lib.rs:20-22, crates/cdivsufsort/src/lib.rs:9-23).
 * Identical signature and return codes: 0 ok; -1 if T==NULL || SA==NULL || n<0;
 * n==0 -> 0; n==1 -> SA[0]=0; n==2 as divsufsort.c:349.  T and SA are HOST
 * pointers; the call copies T to the GPU, builds the SA there (prefix doubling)
 * and copies it back.  Nothing is retained.
 * -------------------------------------------------------------------------- */
int32_t gsa_divsufsort(const uint8_t *T, int32_t *SA, int32_t n);
/* Same, on an explicit CUDA device, optionally returning the per-round log. */
int32_t gsa_divsufsort_ex(const uint8_t *T, int32_t *SA, int32_t n, int32_t device, gsa_build_stats *stats);

/* Device-resident form: d_T (n bytes) and d_SA (n int32) are DEVICE pointers on
 * the current device, `stream` is a cudaStream_t (NULL = default stream).  The
 * workspace is allocated and freed inside the call unless `workspace` is given
 * (gsa_build_workspace_bytes(n) bytes, 256-byte aligned).  Synchronises `stream`
 * before returning. */
size_t gsa_build_workspace_bytes(int32_t n);
int32_t gsa_build_device(const uint8_t *d_T, int32_t *d_SA, int32_t n, void *workspace, size_t workspace_bytes,
                         void *stream, gsa_build_stats *stats);

/* Burrows-Wheeler transform.
 * gsa_divbwt replaces divbwt(T, U, A, n) (c-sources/divsufsort.c:372-405, divsufsort.h): same
 * signature, output convention and return value (the primary index; -1 bad arguments, -2 out
 * of memory; n <= 1 -> n).  `A` is scratch in the reference and is ignored (may be NULL).  HOST
 * pointers.  gsa_bwt_device is the device-pointer form of bw_transform (utils.c:52-110) for a
 * suffix array that is already resident: U[0] = T[n-1], then T[SA[i]-1] for SA[i] != 0 in SA
 * order; *primary_index = slot of suffix 0, plus one.
 * gsa_inverse_bw_transform replaces inverse_bw_transform(T, U, A, n, idx) (utils.c:111-156): same
 * signature, argument checks and return codes (0; -1 for T/U NULL, n < 0, idx < 0, n < idx or
 * n > 0 with idx == 0; -2 out of memory); `A` is ignored.  HOST pointers.  The reference's
 * sequential psi walk is done as a list ranking on the GPU.  (n == 1: U[0] = T[0]; the reference
 * returns without writing U.)  gsa_inverse_bwt_device is the device-pointer form (workspace:
 * gsa_inverse_bwt_workspace_bytes(n) bytes, or NULL to allocate internally). */
int32_t gsa_divbwt(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n);
int32_t gsa_bwt_device(const uint8_t *d_T, const int32_t *d_SA, int32_t n, uint8_t *d_U, int32_t *primary_index,
                       void *stream);
int32_t gsa_inverse_bw_transform(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t idx);
size_t gsa_inverse_bwt_workspace_bytes(int32_t n);
int32_t gsa_inverse_bwt_device(const uint8_t *d_T, uint8_t *d_U, int32_t n, int32_t idx, void *workspace,
                               size_t workspace_bytes, void *stream);

/* Longest-common-prefix array: LCP[0] = 0, LCP[j] = lcp(suffix SA[j-1], suffix SA[j]).
 * The reference has no LCP routine; this is the extension SURVEY.md section 8(f) ranks third
 * (its sa_search, c-sources/utils.c:275-286, carries lmatch/rmatch instead).  Oracle: Kasai's
 * algorithm (oracle/oracle.c).  gsa_lcp_device takes device pointers of a resident text and
 * suffix array (workspace: gsa_lcp_workspace_bytes(n) bytes, or NULL to allocate internally);
 * gsa_lcp takes HOST pointers; gsa_divsufsort_lcp builds SA and LCP in one call (either output
 * may be NULL).  Returns 0, or a negative error code. */
size_t gsa_lcp_workspace_bytes(int32_t n);
int32_t gsa_lcp_device(const uint8_t *d_T, const int32_t *d_SA, int32_t *d_LCP, int32_t n, void *workspace,
                       size_t workspace_bytes, void *stream);
int32_t gsa_lcp(const uint8_t *T, const int32_t *SA, int32_t *LCP, int32_t n, int32_t device);
int32_t gsa_divsufsort_lcp(const uint8_t *T, int32_t *SA, int32_t *LCP, int32_t n, int32_t device);

/* O(n) validity check of a suffix array on the GPU (device pointers).
 * Replaces sacabase::verify (crates/sacabase/src/lib.rs:127-149) / sufcheck
 * (c-sources/utils.c:160-241) at sizes where the O(n * LCP) pairwise check is
 * infeasible.  Returns 0 if d_SA is the suffix array of d_T, 1 if not
 * (*bad_index receives one offending SA slot), negative on error. */
int32_t gsa_sufcheck_device(const uint8_t *d_T, const int32_t *d_SA, int32_t n, void *stream, int64_t *bad_index);
/* Host-pointer convenience form (copies in, checks on `device`). */
int32_t gsa_sufcheck(const uint8_t *T, const int32_t *SA, int32_t n, int32_t device, int64_t *bad_index);

/* ----------------------------------------------------------------------------
 * Resident index = text + suffix array kept in HBM for queries.
 * Replaces sacabase::SuffixArray<'a, i32> (crates/sacabase/src/lib.rs:152-188):
 *   gsa_index_create      <-> divsufsort::sort(text)            (lib.rs:25-29)
 *   gsa_index_from_parts  <-> SuffixArray::new(text, sa)        (sacabase :170-172)
 *   gsa_index_sa          <-> SuffixArray::into_parts           (:175-177)
 *   gsa_index_verify      <-> SuffixArray::verify               (:180-182)
 * `T` is a host pointer; it is copied, not retained.
 * -------------------------------------------------------------------------- */
typedef struct gsa_index gsa_index;
int32_t gsa_index_create(const uint8_t *T, int64_t n, int32_t device, gsa_index **out, gsa_build_stats *stats);
/* sa_len must equal n (GSA_EINVAL otherwise: the handle keeps one length for both arrays), and every
 * entry of SA must be a text position (checked on the device; GSA_EPANIC otherwise -- the reference
 * panics on its slice bounds check the first time such an entry is used, lib.rs:53-57). */
int32_t gsa_index_from_parts(const uint8_t *T, int64_t n, const int32_t *SA, int64_t sa_len, int32_t device,
                             gsa_index **out);
int64_t gsa_index_len(const gsa_index *ix);
int32_t gsa_index_device(const gsa_index *ix);
int32_t gsa_index_sa(const gsa_index *ix, int32_t *out_sa /* host, n entries */);
int32_t gsa_index_verify(const gsa_index *ix, int64_t *bad_index);
const uint8_t *gsa_index_device_text(const gsa_index *ix);
const int32_t *gsa_index_device_sa(const gsa_index *ix);
void gsa_index_destroy(gsa_index *ix);

/* ----------------------------------------------------------------------------
 * Batched search.  Patterns are concatenated in `pats`; pattern q is
 * pats[pat_off[q] .. pat_off[q+1]).  All pointers are HOST pointers.
 *
 * gsa_lsm_batch replaces StringIndex::longest_substring_match
 *   (crates/sacabase/src/lib.rs:39-99,160-163,190-197) applied to each pattern:
 *   the exact narrowing rule (mid = len/2; needle > suff(mid) ? [mid..] : [..=mid]),
 *   the len-1 / len-2 terminal cases and the `x > y` tie-break are reproduced, so
 *   (start, len) equal the reference's for every input.  Returns GSA_EPANIC for an
 *   empty index (the reference panics, lib.rs:89-91).
 *
 * gsa_search_all_batch replaces sa_search (c-sources/utils.c:258-325,
 *   divsufsort.h:152-157), the in-tree semantics behind the `search_all` name:
 *   out_count[q] = number of occurrences, out_left[q] = *idx (first SA slot of the
 *   occurrence range, or the insertion point on a miss; empty pattern -> count n,
 *   left 0; empty text -> count 0, left -1).
 * gsa_contains_batch: out[q] = (count > 0).
 * -------------------------------------------------------------------------- */
int32_t gsa_lsm_batch(const gsa_index *ix, const uint8_t *pats, const uint64_t *pat_off, uint64_t Q,
                      uint64_t *out_start, uint32_t *out_len);
int32_t gsa_search_all_batch(const gsa_index *ix, const uint8_t *pats, const uint64_t *pat_off, uint64_t Q,
                             int32_t *out_left, int32_t *out_count);
int32_t gsa_contains_batch(const gsa_index *ix, const uint8_t *pats, const uint64_t *pat_off, uint64_t Q,
                           uint8_t *out);

/* Device-resident forms (all pointers are device pointers on the index's device;
 * results stay on the device; asynchronous on `stream`).  `max_pat_len` is the longest
 * pattern of the batch, or 0 if unknown; it only selects how many lanes share a pattern
 * (8 lanes for <= 32 bytes, 16 for <= 64, else 32) -- results do not depend on it.  For gsa_lsm_device the partition parameters implement
 * sacapart's per-shard step (crates/sacapart/src/lib.rs:71-92): `offset` is added
 * to start, and when `accumulate` != 0 a result only replaces
 * (io_start[q], io_len[q]) if its len is strictly greater. */
int32_t gsa_lsm_device(const gsa_index *ix, const uint8_t *d_pats, const uint64_t *d_pat_off, uint64_t Q,
                       uint32_t max_pat_len, uint64_t offset, int32_t accumulate, uint64_t *d_io_start,
                       uint32_t *d_io_len, void *stream);
int32_t gsa_search_all_device(const gsa_index *ix, const uint8_t *d_pats, const uint64_t *d_pat_off, uint64_t Q,
                              uint32_t max_pat_len, int32_t *d_out_left, int32_t *d_out_count, void *stream);

/* Merge step of a fanned-out partitioned query: for every q keep the longer match;
 * on equal length keep the one with the smaller start (= lower partition index,
 * because partitions are disjoint ascending ranges; sacapart lib.rs:86-92).
 * d_start/d_len hold `nsets` result sets of Q entries each (set-major); the winner
 * is written to set 0.  Device pointers. */
int32_t gsa_lsm_reduce_device(uint64_t *d_start, uint32_t *d_len, uint64_t Q, uint32_t nsets, void *stream);

/* ----------------------------------------------------------------------------
 * Partitioned suffix array.
 * Replaces sacapart::PartitionedSuffixArray (crates/sacapart/src/lib.rs:26-98):
 *   gsa_part_create          <-> ::new(text, num_partitions, divsufsort::sort)  (:39-58)
 *        chunk = n / num_partitions + 1 bytes (:43); shard i is built on
 *        devices[i % ndev]; shards on different devices are built concurrently
 *        (the rayon par_chunks of :45-49 becomes one host thread per device).
 *   gsa_part_num_partitions  <-> ::num_partitions()                             (:60-62)
 *   gsa_part_lsm_batch       <-> StringIndex::longest_substring_match           (:69-97)
 *        offset / may_extend / strict-greater replacement reproduced exactly.
 * `T` (host) must stay valid for the lifetime of the handle, as the reference's
 * borrow `&'a [u8]` requires (:31); it is used to top up the per-shard halo that the
 * may_extend rule (:77-84) reads past the shard end.
 * Queries on one handle are serialised internally (a query may grow a shard's halo, which replaces
 * its device text buffer); different handles are independent.
 * num_partitions == 0 -> GSA_EPANIC (the reference divides by zero, :43);
 * querying a handle with zero partitions (empty text) -> GSA_EPANIC (:94-96).
 * -------------------------------------------------------------------------- */
typedef struct gsa_part gsa_part;
int32_t gsa_part_create(const uint8_t *T, uint64_t n, uint64_t num_partitions, const int32_t *devices, int32_t ndev,
                        gsa_part **out);
uint64_t gsa_part_num_partitions(const gsa_part *p);
uint64_t gsa_part_partition_size(const gsa_part *p);
const gsa_index *gsa_part_shard(const gsa_part *p, uint64_t i);
int32_t gsa_part_lsm_batch(gsa_part *p, const uint8_t *pats, const uint64_t *pat_off, uint64_t Q, uint64_t *out_start,
                           uint32_t *out_len);
void gsa_part_destroy(gsa_part *p);

/* Single shard with halo, for one-process-per-GPU deployments (torch.distributed
 * ranks each own the shards i with i % world == rank): builds the SA of
 * T[offset .. offset+len) and keeps `halo` further bytes of T behind it so that
 * gsa_lsm_device can apply the may_extend rule locally. */
int32_t gsa_index_create_shard(const uint8_t *T_full, uint64_t n_full, uint64_t offset, uint64_t len, uint64_t halo,
                               int32_t device, gsa_index **out, gsa_build_stats *stats);

/* ----------------------------------------------------------------------------
 * Utilities.
 * -------------------------------------------------------------------------- */
/* Pinned host memory for the end-to-end path (plain cudaHostAlloc / cudaFreeHost). */
void *gsa_host_alloc(size_t bytes);
void gsa_host_free(void *p);
/* The host-pointer entry points keep one scratch allocation per device between calls
 * (device text + SA + sort workspace; allocating tens of GB per call would dominate
 * end-to-end time).  This returns all idle cached blocks to the CUDA driver. */
void gsa_release_cached_memory(void);
/* Thread-local text of the last CUDA/runtime error seen by this thread. */
const char *gsa_last_error(void);
/* Library version string, also names the compiled arch ("sm_100a"). */
const char *gsa_version(void);
/* Number of visible CUDA devices (<= 0: no usable GPU). */
int32_t gsa_device_count(void);
/* The calling thread's current CUDA device (what gsa_divsufsort() builds on); 0 if there is none. */
int32_t gsa_current_device(void);

#ifdef __cplusplus
}
#endif
#endif /* GSA_H */
