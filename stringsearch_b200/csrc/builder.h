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
struct TextView {
  const u8 *text;
  const i32 *sa;
  u64 n;
  u64 text_avail;
};
// max_pat_len: longest pattern of the batch, or 0 if unknown (only picks the lanes-per-pattern variant)
int lsm_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, u32 max_pat_len, u64 offset,
               int accumulate, u64 *d_io_start, u32 *d_io_len, cudaStream_t st);
int search_all_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, u32 max_pat_len, i32 *d_left,
                      i32 *d_count, cudaStream_t st);
int lsm_reduce_device(u64 *d_start, u32 *d_len, u64 Q, u32 nsets, cudaStream_t st);

}  // namespace gsa
