// bwt.cu -- Burrows-Wheeler transform from the suffix array.
//
// Replaces libdivsufsort's divbwt / bw_transform (reference:
// crates/cdivsufsort/c-sources/divsufsort.c:372-405, utils.c:52-110, header divsufsort.h:78-127)
// with the same output convention:
//     U[0] = T[n-1];  the other n-1 characters are T[SA[i]-1] for the slots i with SA[i] != 0,
//     in SA order;    primary index = (slot holding suffix 0) + 1.
// One gather kernel: a streaming read of SA, a random 1-byte read of T per slot.
//
// Inverse transform (inverse_bw_transform, utils.c:111-156): the reference follows the psi
// permutation B -- a stable counting sort of the transformed string -- from row `idx`, one
// dependent step per output byte.  Here the same walk is a list ranking (Helman & JaJa, 1999):
//   stable_byte_order_device   B' = stable order of the bytes (one radix pass)
//   k_ibwt_links               next row of every row; every STRIDE-th row and the start row are
//                              splitters (marked in bit 31 of their link)
//   k_ibwt_walk1               one thread per splitter: length of its sublist, next splitter
//   k_ibwt_jump  x log2(S)     pointer jumping over the S splitters: bytes from each one to the end
//   k_ibwt_walk2               one thread per splitter: walk again, writing the output bytes
// Work O(n): two random 4-byte reads per output byte.
#include "builder.h"

namespace gsa {

__global__ void __launch_bounds__(256) k_find_zero(const i32 *__restrict__ SA, u32 n, u32 *__restrict__ i0) {
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (SA[i] == 0) *i0 = i;
}

__global__ void __launch_bounds__(256) k_bwt(const u8 *__restrict__ T, const i32 *__restrict__ SA, u32 n,
                                             const u32 *__restrict__ i0p, u8 *__restrict__ U) {
  const u32 i0 = *i0p;
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (i == i0) continue;
    const u32 s = (u32)SA[i];
    U[i < i0 ? i + 1 : i] = __ldg(T + s - 1);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) U[0] = T[n - 1];
}

int bwt_device(const u8 *d_T, const i32 *d_SA, u32 n, u8 *d_U, i32 *primary_index, cudaStream_t st) {
  if (n <= 1) {
    if (n == 1) GSA_TRY(cudaMemcpyAsync(d_U, d_T, 1, cudaMemcpyDeviceToDevice, st));
    if (primary_index) *primary_index = (i32)n;  // divsufsort.c:380
    GSA_TRY(cudaStreamSynchronize(st));
    return GSA_OK;
  }
  u32 *d_i0 = nullptr;
  GSA_TRY(cudaMalloc(&d_i0, sizeof(u32)));
  struct Free { u32 *p; ~Free() { cudaFree(p); } } guard{d_i0};
  int dev = 0, sms = kDefaultSMs;
  GSA_TRY(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const u32 blocks = (u32)std::min<u64>((u64)sms * 8, div_up(n, 256));
  k_find_zero<<<blocks, 256, 0, st>>>(d_SA, n, d_i0);
  GSA_TRY(cudaGetLastError());
  k_bwt<<<blocks, 256, 0, st>>>(d_T, d_SA, n, d_i0, d_U);
  GSA_TRY(cudaGetLastError());
  u32 i0 = 0;
  GSA_TRY(cudaMemcpyAsync(&i0, d_i0, sizeof(u32), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaStreamSynchronize(st));
  if (primary_index) *primary_index = (i32)(i0 + 1);
  return GSA_OK;
}

// ------------------------------------------------------------------------------------
// Inverse BWT
// ------------------------------------------------------------------------------------
namespace {

constexpr u32 IB_STRIDE = 1024;         // one splitter per this many rows
constexpr u32 IB_MARK = 0x80000000u;    // link word: this row is a splitter
constexpr u32 IB_END = 0x7fffffffu;     // link value: the walk ends after this row
constexpr u32 IB_NONE = 0xffffffffu;    // splitter list: no successor

// order[q] (index into the transformed string) -> link[q]: 0-based row that follows row q.
// utils.c:143-144: B[..] = i for i < idx, i + 1 otherwise (1-based rows); the walk goes p = B[p - 1].
__global__ void __launch_bounds__(256) k_ibwt_links(u32 *__restrict__ link, u32 n, u32 idx, u32 q0) {
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const u32 i = link[q];
    const u32 nxt = (i < idx) ? (i == 0u ? IB_END : i - 1u) : i;
    link[q] = nxt | ((q % IB_STRIDE == 0u || q == q0) ? IB_MARK : 0u);
  }
}

__device__ __forceinline__ u32 splitter_row(u32 id, u32 S, u32 q0) { return id < S ? id * IB_STRIDE : q0; }
__device__ __forceinline__ u32 splitter_id(u32 q, u32 S, u32 q0) { return (q % IB_STRIDE == 0u) ? q / IB_STRIDE : S; }

__global__ void __launch_bounds__(256) k_ibwt_walk1(const u32 *__restrict__ link, u32 n, u32 S, u32 nsplit, u32 q0,
                                                    u32 *__restrict__ len, u32 *__restrict__ succ) {
  const u32 id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nsplit) return;
  u32 l = 1;
  u32 nxt = __ldg(link + splitter_row(id, S, q0)) & ~IB_MARK;
  while (nxt != IB_END && l <= n) {
    const u32 w = __ldg(link + nxt);
    if (w & IB_MARK) break;
    ++l;
    nxt = w & ~IB_MARK;
  }
  len[id] = l;
  succ[id] = (nxt == IB_END || l > n) ? IB_NONE : splitter_id(nxt, S, q0);
}

// one pointer-jumping round: dist = bytes from the start of this sublist to the end of the text
__global__ void __launch_bounds__(256) k_ibwt_jump(const u32 *__restrict__ dist_in, const u32 *__restrict__ succ_in,
                                                   u32 *__restrict__ dist_out, u32 *__restrict__ succ_out, u32 nsplit) {
  const u32 id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nsplit) return;
  const u32 s = succ_in[id];
  u32 d = dist_in[id], s2 = IB_NONE;
  if (s != IB_NONE) {
    d += dist_in[s];
    s2 = succ_in[s];
  }
  dist_out[id] = d;
  succ_out[id] = s2;
}

__global__ void __launch_bounds__(256) k_ibwt_walk2(const u32 *__restrict__ link, const u32 *__restrict__ counts, u32 n,
                                                    u32 S, u32 nsplit, u32 q0, const u32 *__restrict__ len,
                                                    const u32 *__restrict__ dist, u8 *__restrict__ U) {
  // first-column byte of row q = the c with C[c] <= q < C[c + 1]  (utils.c:147 binarysearch_lower)
  __shared__ u32 s_c[257];
  if (threadIdx.x == 0) {
    u32 acc = 0;
    for (int c = 0; c < 256; ++c) { s_c[c] = acc; acc += counts[c]; }
    s_c[256] = acc;
  }
  __syncthreads();
  const u32 id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nsplit) return;
  const u32 d = dist[id], l = len[id];
  if (d > n || l > d) return;  // not a transform of any text: leave the output alone rather than run wild
  u32 pos = n - d;
  u32 q = splitter_row(id, S, q0);
  for (u32 k = 0; k < l; ++k) {
    u32 lo = 0, hi = 256;  // s_c[lo] <= q < s_c[hi]
    while (hi - lo > 1u) {
      const u32 mid = (lo + hi) >> 1;
      if (s_c[mid] <= q) lo = mid; else hi = mid;
    }
    U[pos + k] = (u8)lo;
    q = __ldg(link + q) & ~IB_MARK;
  }
}

struct IbwtLayout { u32 *link, *counts, *len, *succ[2], *dist[2]; char *sort_ws; size_t sort_bytes; u32 S; size_t total; };
IbwtLayout ibwt_layout(char *base, u32 n) {
  IbwtLayout y;
  Carve c{base, 0};
  y.S = (u32)div_up(n, IB_STRIDE);
  y.link = c.take<u32>(n);
  y.counts = c.take<u32>(256);
  y.len = c.take<u32>(y.S + 1);
  for (int i = 0; i < 2; ++i) { y.succ[i] = c.take<u32>(y.S + 1); y.dist[i] = c.take<u32>(y.S + 1); }
  y.sort_bytes = byte_order_workspace_bytes(n);
  y.sort_ws = c.take<char>(y.sort_bytes);
  y.total = c.used;
  return y;
}

}  // namespace

size_t inverse_bwt_workspace_bytes(u32 n) { return ibwt_layout(nullptr, n == 0 ? 1 : n).total + 256; }

// d_T: transformed string (n bytes), idx: primary index (1..n) as returned by divbwt; d_U: output.
int inverse_bwt_device(const u8 *d_T, u8 *d_U, u32 n, u32 idx, void *workspace, size_t workspace_bytes, cudaStream_t st) {
  if (n == 0) return GSA_OK;
  if (n == 1) {
    GSA_TRY(cudaMemcpyAsync(d_U, d_T, 1, cudaMemcpyDeviceToDevice, st));
    GSA_TRY(cudaStreamSynchronize(st));
    return GSA_OK;
  }
  char *owned = nullptr;
  const size_t need = inverse_bwt_workspace_bytes(n);
  if (workspace == nullptr) {
    cudaError_t e = cudaMalloc(&owned, need);
    if (e != cudaSuccess) {
      set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      cudaGetLastError();
      return GSA_ENOMEM;
    }
    workspace = owned;
  } else if (workspace_bytes < need) {
    set_error("workspace too small", __FILE__, __LINE__);
    return GSA_EINVAL;
  }
  struct Free { char *p; ~Free() { if (p) cudaFree(p); } } guard{owned};
  const size_t mis = (256 - (reinterpret_cast<uintptr_t>(workspace) & 255)) & 255;
  const IbwtLayout y = ibwt_layout(static_cast<char *>(workspace) + mis, n);
  int dev = 0, sms = kDefaultSMs;
  GSA_TRY(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  GSA_TRY_RC(stable_byte_order_device(d_T, n, y.link, y.counts, y.sort_ws, y.sort_bytes, st));
  const u32 q0 = idx - 1u;  // the walk starts at row idx (1-based)
  const u32 S = y.S;
  const u32 nsplit = (q0 % IB_STRIDE == 0u) ? S : S + 1u;
  const u32 blocks = (u32)std::min<u64>((u64)sms * 8, div_up(n, 256));
  k_ibwt_links<<<blocks, 256, 0, st>>>(y.link, n, idx, q0);
  GSA_TRY(cudaGetLastError());
  const u32 sblocks = (u32)div_up(nsplit, 256);
  k_ibwt_walk1<<<sblocks, 256, 0, st>>>(y.link, n, S, nsplit, q0, y.len, y.succ[0]);
  GSA_TRY(cudaGetLastError());
  GSA_TRY(cudaMemcpyAsync(y.dist[0], y.len, (size_t)nsplit * sizeof(u32), cudaMemcpyDeviceToDevice, st));
  int cur = 0;
  for (u32 span = 1; span < nsplit; span <<= 1) {
    k_ibwt_jump<<<sblocks, 256, 0, st>>>(y.dist[cur], y.succ[cur], y.dist[cur ^ 1], y.succ[cur ^ 1], nsplit);
    GSA_TRY(cudaGetLastError());
    cur ^= 1;
  }
  k_ibwt_walk2<<<sblocks, 256, 0, st>>>(y.link, y.counts, n, S, nsplit, q0, y.len, y.dist[cur], d_U);
  GSA_TRY(cudaGetLastError());
  GSA_TRY(cudaStreamSynchronize(st));
  return GSA_OK;
}

}  // namespace gsa
