// builder.h -- internal (non-ABI) interfaces between the .cu translation units.
#pragma once
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace gsa {

// sa_build.cu
size_t build_workspace_bytes(u32 n);
int build_sa_device(const u8 *d_T, i32 *d_SA, u32 n, void *workspace, size_t workspace_bytes, cudaStream_t st,
                    gsa_build_stats *stats);

// Stable order of a byte string (one radix pass): order[q] = index of the q-th byte in (value, index)
// order; counts[256] (device, optional) receives the byte histogram.
size_t byte_order_workspace_bytes(u32 n);
int stable_byte_order_device(const u8 *d_bytes, u32 n, u32 *d_order, u32 *d_counts, void *workspace, size_t workspace_bytes,
                             cudaStream_t st);

// verify.cu
size_t sufcheck_workspace_bytes(u32 n);
int sufcheck_device(const u8 *d_T, const i32 *d_SA, u32 n, cudaStream_t st, i64 *bad_index);
// *bad_slot = first slot whose entry is not in [0, n), or -1.  Synchronises `st`.
int sa_range_check_device(const i32 *d_SA, u32 n, cudaStream_t st, i64 *bad_slot);

// bwt.cu
int bwt_device(const u8 *d_T, const i32 *d_SA, u32 n, u8 *d_U, i32 *primary_index, cudaStream_t st);
size_t inverse_bwt_workspace_bytes(u32 n);
int inverse_bwt_device(const u8 *d_T, u8 *d_U, u32 n, u32 idx, void *workspace, size_t workspace_bytes, cudaStream_t st);

// lcp.cu
size_t lcp_workspace_bytes(u32 n);
int lcp_device(const u8 *d_T, const i32 *d_SA, u32 n, i32 *d_LCP, void *workspace, size_t workspace_bytes,
               cudaStream_t st);

// search.cu
// Text view of an index: the suffix array covers text[0, n); text_avail >= n bytes are
// readable (halo behind a shard, used only by the may_extend rule).
// Prefix-bucket table of an index (search.cu, "top of the tree"): with code[] the dense code of every byte
// value (number of smaller byte values that occur in the text; b bits per code) and
//   key(S) = the first k symbols of S as a k*b-bit number, zero padded when S is shorter,
// T[c] = number of suffixes whose key is below c (c = 0 .. 2^(k b)).  Keys are non-decreasing along the
// suffix array, so the insertion point of a pattern lies in [T[key], T[key + 1]] and a search starts there
// instead of at [0, n].  k == 0: no table (tiny texts, GSA_NO_ACCEL).
struct AccelView {
  const u32 *T;
  u32 k, b;
  u32 present[8];  // bitmap of the byte values that occur
  u8 code[256];
};
struct TextView {
  const u8 *text;
  const i32 *sa;
  u64 n;
  u64 text_avail;
  AccelView ac;
};
// Builds the table for (d_T[0, n), d_SA) on the current device; *d_table is cudaMalloc'd (caller frees).
// bits = upper bound on k * b (0: default).  Synchronises `st`.
int accel_build_device(const u8 *d_T, const i32 *d_SA, u64 n, u32 bits, AccelView *out, u32 **d_table, cudaStream_t st);
// max_pat_len: longest pattern of the batch, or 0 if unknown (only picks the lanes-per-pattern variant)
int lsm_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, u32 max_pat_len, u64 offset,
               int accumulate, u64 *d_io_start, u32 *d_io_len, cudaStream_t st);
int search_all_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, u32 max_pat_len, i32 *d_left,
                      i32 *d_count, cudaStream_t st);
int lsm_reduce_device(u64 *d_start, u32 *d_len, u64 Q, u32 nsets, cudaStream_t st);

}  // namespace gsa
