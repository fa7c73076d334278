// common.cuh -- shared types, error plumbing and warp-level helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsa.h"

namespace gsa {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

// Thread-local last-error text (api.cu owns the storage).
void set_error(const char *what, const char *file, int line);

#define GSA_TRY(expr)                                                   \
  do {                                                                  \
    cudaError_t e__ = (expr);                                           \
    if (e__ != cudaSuccess) {                                           \
      ::gsa::set_error(cudaGetErrorString(e__), __FILE__, __LINE__);    \
      return GSA_ECUDA;                                                 \
    }                                                                   \
  } while (0)

#define GSA_TRY_RC(expr)                 \
  do {                                   \
    int rc__ = (expr);                   \
    if (rc__ != GSA_OK) return rc__;     \
  } while (0)

static inline u64 div_up(u64 a, u64 b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
// Carves 256-byte aligned arrays out of one workspace block (p may be null: size computation only).
struct Carve {
  char *p;
  size_t used;
  template <typename T>
  T *take(size_t count) {
    T *r = reinterpret_cast<T *>(p + used);
    used += align_up(count * sizeof(T), 256);
    return r;
  }
};
// bits needed to represent values 0..v  (bits_for(0) = 1 so a key is never 0 bits wide)
static inline u32 bits_for(u64 v) {
  u32 b = 1;
  while (b < 64 && (v >> b) != 0) ++b;
  return b;
}

// B200: 148 SMs. Queried once per device at run time; this is only the fallback.
constexpr int kDefaultSMs = 148;

#ifdef __CUDACC__
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt() {
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ u32 ld_volatile_u32(const u32 *p) {
  u32 v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(u32 *p, u32 v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ld_volatile_u64(const u64 *p) {
  u64 v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(u64 *p, u64 v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Streaming (read-once) loads: do not allocate in L1.
__device__ __forceinline__ u64 ld_stream_u64(const u64 *p) {
  u64 v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ u32 ld_stream_u32(const u32 *p) {
  u32 v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u128(const void *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
#endif  // __CUDACC__

}  // namespace gsa
