// bwt.cu -- Burrows-Wheeler transform from the suffix array.
//
// Replaces libdivsufsort's divbwt / bw_transform (reference:
// crates/cdivsufsort/c-sources/divsufsort.c:372-405, utils.c:52-110, header divsufsort.h:78-127)
// with the same output convention:
//     U[0] = T[n-1];  the other n-1 characters are T[SA[i]-1] for the slots i with SA[i] != 0,
//     in SA order;    primary index = (slot holding suffix 0) + 1.
// One gather kernel: a streaming read of SA, a random 1-byte read of T per slot.
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

}  // namespace gsa
