// verify.cu -- O(n) suffix-array validity check on the GPU.
//
// Stands in for sacabase::verify (reference: crates/sacabase/src/lib.rs:127-149, pairwise
// suffix comparison, O(n * LCP)) and libdivsufsort's sufcheck
// (crates/cdivsufsort/c-sources/utils.c:160-241) at sizes where neither is practical.
// SA is the suffix array of T iff
//   (1) SA is a permutation of 0..n-1, and
//   (2) for every adjacent pair a = SA[j-1], b = SA[j]:  T[a] < T[b], or T[a] == T[b] and
//       ISA[a+1] < ISA[b+1]  with ISA[n] = -1 (the empty suffix sorts first).
// (2) holding for all adjacent pairs implies, by induction on the suffix length, that the
// whole order is the lexicographic one.
#include "builder.h"

namespace gsa {

__global__ void __launch_bounds__(256) k_isa_scatter(const i32 *__restrict__ SA, u32 n, u32 *__restrict__ isa,
                                                     unsigned long long *__restrict__ bad) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const u32 s = (u32)SA[j];
  if (s >= n) {
    atomicMin(bad, (unsigned long long)j);
    return;
  }
  isa[s] = j;
}

__global__ void __launch_bounds__(256) k_sufcheck(const u8 *__restrict__ T, const i32 *__restrict__ SA, u32 n,
                                                  const u32 *__restrict__ isa, unsigned long long *__restrict__ bad) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const u32 b = (u32)SA[j];
  if (b >= n) return;  // already reported
  if (isa[b] != j) {   // duplicate entry somewhere: not a permutation
    atomicMin(bad, (unsigned long long)j);
    return;
  }
  if (j == 0) return;
  const u32 a = (u32)SA[j - 1];
  if (a >= n) return;
  const u8 ca = T[a], cb = T[b];
  bool ok;
  if (ca != cb) {
    ok = ca < cb;
  } else {
    const i64 ra = (a + 1 < n) ? (i64)isa[a + 1] : -1;
    const i64 rb = (b + 1 < n) ? (i64)isa[b + 1] : -1;
    ok = ra < rb;
  }
  if (!ok) atomicMin(bad, (unsigned long long)(j - 1));
}

// Failure path only: the reference's verify (sacabase lib.rs:143-147) reports the FIRST adjacent pair
// that is not in increasing order, found by comparing the suffixes themselves.  The rank criterion
// above decides pass / fail in O(n), but the pair it flags can be one whose own order is fine (its
// successors' ranks are what is inverted), so after a failure the pairs are compared directly, in
// ascending order, stopping at the smallest offender found so far.  A pair that agrees on
// kDirectCap bytes is decided by the ranks of the suffixes behind those bytes.
constexpr u32 kDirectCap = 1u << 16;

__global__ void __launch_bounds__(256) k_first_misordered(const u8 *__restrict__ T, const i32 *__restrict__ SA, u32 n,
                                                          const u32 *__restrict__ isa, unsigned long long *__restrict__ first) {
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j + 1 < n; j += stride) {
    if ((unsigned long long)j >= *reinterpret_cast<volatile unsigned long long *>(first)) return;
    const u32 a = (u32)SA[j], b = (u32)SA[j + 1];
    if (a >= n || b >= n) continue;
    bool bad = false, decided = false;
    if (a == b) { bad = true; decided = true; }
    for (u32 k = 0; !decided && k < kDirectCap; ++k) {
      if (a + k >= n) { decided = true; break; }               // suffix a is a proper prefix of suffix b: in order
      if (b + k >= n) { bad = true; decided = true; break; }   // the other way round
      const u8 ca = T[a + k], cb = T[b + k];
      if (ca != cb) { bad = ca > cb; decided = true; }
    }
    if (!decided) bad = isa[a + kDirectCap] > isa[b + kDirectCap];
    if (bad) atomicMin(first, (unsigned long long)j);
  }
}

__global__ void __launch_bounds__(256) k_sa_range(const i32 *__restrict__ SA, u32 n, unsigned long long *__restrict__ bad) {
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    if ((u32)SA[j] >= n) atomicMin(bad, (unsigned long long)j);
}

int sa_range_check_device(const i32 *d_SA, u32 n, cudaStream_t st, i64 *bad_slot) {
  *bad_slot = -1;
  if (n == 0) return GSA_OK;
  unsigned long long *bad = nullptr;
  cudaError_t e = cudaMalloc(&bad, sizeof(unsigned long long));
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e), __FILE__, __LINE__);
    cudaGetLastError();
    return GSA_ENOMEM;
  }
  struct Free { void *p; ~Free() { cudaFree(p); } } guard{bad};
  GSA_TRY(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), st));
  k_sa_range<<<(u32)std::min<u64>(div_up(n, 256), 148 * 16), 256, 0, st>>>(d_SA, n, bad);
  GSA_TRY(cudaGetLastError());
  unsigned long long h = 0;
  GSA_TRY(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaStreamSynchronize(st));
  if (h != ~0ull) *bad_slot = (i64)h;
  return GSA_OK;
}

size_t sufcheck_workspace_bytes(u32 n) { return align_up((size_t)n * 4, 256) + 256; }

int sufcheck_device(const u8 *d_T, const i32 *d_SA, u32 n, cudaStream_t st, i64 *bad_index) {
  if (bad_index) *bad_index = -1;
  if (n == 0) return 0;
  char *ws = nullptr;
  cudaError_t e = cudaMalloc(&ws, sufcheck_workspace_bytes(n));
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e), __FILE__, __LINE__);
    cudaGetLastError();
    return GSA_ENOMEM;
  }
  struct Free { char *p; ~Free() { cudaFree(p); } } guard{ws};
  u32 *isa = reinterpret_cast<u32 *>(ws);
  unsigned long long *bad = reinterpret_cast<unsigned long long *>(ws + align_up((size_t)n * 4, 256));
  GSA_TRY(cudaMemsetAsync(isa, 0xff, (size_t)n * 4, st));
  GSA_TRY(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), st));
  const u32 blocks = (u32)div_up(n, 256);
  k_isa_scatter<<<blocks, 256, 0, st>>>(d_SA, n, isa, bad);
  GSA_TRY(cudaGetLastError());
  k_sufcheck<<<blocks, 256, 0, st>>>(d_T, d_SA, n, isa, bad);
  GSA_TRY(cudaGetLastError());
  unsigned long long h = 0;
  GSA_TRY(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaStreamSynchronize(st));
  if (h == ~0ull) return 0;
  if (bad_index) {
    // which pair would the reference name?  (only worth the extra pass when somebody asks)
    unsigned long long *first = bad;
    GSA_TRY(cudaMemsetAsync(first, 0xff, sizeof(unsigned long long), st));
    k_first_misordered<<<(u32)std::min<u64>(blocks, 148 * 8), 256, 0, st>>>(d_T, d_SA, n, isa, first);
    GSA_TRY(cudaGetLastError());
    unsigned long long f = 0;
    GSA_TRY(cudaMemcpyAsync(&f, first, sizeof(f), cudaMemcpyDeviceToHost, st));
    GSA_TRY(cudaStreamSynchronize(st));
    *bad_index = (i64)(f != ~0ull ? f : h);
  }
  return 1;
}

}  // namespace gsa
