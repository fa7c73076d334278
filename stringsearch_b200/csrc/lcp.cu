// lcp.cu -- longest-common-prefix array from text + suffix array.
//
//     LCP[0] = 0,   LCP[j] = length of the longest common prefix of suffixes SA[j-1] and SA[j].
//
// The reference has no LCP routine (libdivsufsort's sa_search carries lmatch/rmatch instead,
// utils.c:275-286); the oracle for this file is Kasai's algorithm in oracle/oracle.c.
//
// Kasai's algorithm is one sequential walk over the text.  The parallel form used here rests
// on the same lemma, applied per position instead of along the walk
// (Karkkainen, Manzini, Puglisi: "Permuted longest-common-prefix array", CPM 2009):
//   phi[i]  = the suffix preceding suffix i in the SA        (phi[SA[j]] = SA[j-1])
//   PLCP[i] = lcp(i, phi[i])                                  (LCP[j] = PLCP[SA[j]])
//   if T[i-1] == T[phi[i]-1] then PLCP[i] = PLCP[i-1] - 1     ("reducible")
//   the PLCP values of the other ("irreducible") positions sum to at most 2 n log2 n.
// So: compare text only at the irreducible positions (bounded total work whatever the text),
// and fill the reducible ones by a prefix maximum of reach[i] = i + PLCP[i], which is
// non-decreasing in i (PLCP[i] >= PLCP[i-1] - 1).
//
//   k_phi            scatter  phi[SA[j]] = SA[j-1]                       (4 B random write / suffix)
//   k_irreducible    one thread per position: classify; irreducible -> compare up to 64 bytes,
//                    longer matches are left to
//   k_long           one warp per match, 2 KiB per step
//   k_reach_tiles / k_reach_spine / k_reach_apply   inclusive prefix maximum, PLCP in place
//   k_lcp_gather     LCP[j] = PLCP[SA[j]]                                (4 B random read / suffix)
#include "builder.h"

namespace gsa {
namespace {

constexpr u32 PHI_NONE = 0xffffffffu;  // the suffix at SA[0] has no predecessor
constexpr u32 REACH_LONG = 0xffffffffu;  // marker: match longer than SHORT_BYTES, finished by k_long
constexpr u32 SHORT_BYTES = 64;
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

// nb (1..8) bytes at an arbitrary address as a little-endian word, from aligned 8-byte loads;
// only words holding at least one requested byte are touched (no padding behind T needed).
__device__ __forceinline__ u64 load8(const u8 *p, u32 nb) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const u64 *w = reinterpret_cast<const u64 *>(a & ~(uintptr_t)7);
  const u32 mis = (u32)(a & 7u);
  const u64 lo = __ldg(w);
  if (mis == 0u) return lo;
  u64 v = lo >> (8u * mis);
  if (mis + nb > 8u) v |= __ldg(w + 1) << (64u - 8u * mis);
  return v;
}

// number of equal leading bytes of the two words, at most nb
__device__ __forceinline__ u32 equal_bytes(u64 x, u64 y, u32 nb) {
  const u64 d = x ^ y;
  const u32 e = d ? ((u32)(__ffsll((long long)d) - 1) >> 3) : 8u;
  return min(e, nb);
}

__global__ void __launch_bounds__(256) k_phi(const i32 *__restrict__ SA, u32 n, u32 *__restrict__ phi) {
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    phi[(u32)SA[j]] = j ? (u32)SA[j - 1] : PHI_NONE;
}

// in: a[i] = phi[i].  out: a[i] = i + PLCP[i] for irreducible positions (REACH_LONG if the
// comparison is unfinished: then aux[i] = phi[i]), 0 for reducible ones.
__global__ void __launch_bounds__(256) k_irreducible(const u8 *__restrict__ T, u32 n, u32 *__restrict__ a,
                                                     u32 *__restrict__ aux, u32 *__restrict__ n_long) {
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u32 j = a[i];
    u32 out;
    if (j == PHI_NONE) {
      out = i;  // PLCP = 0
    } else if (i > 0 && j > 0 && __ldg(T + i - 1) == __ldg(T + j - 1)) {
      out = 0;  // reducible: PLCP[i] = PLCP[i-1] - 1, filled by the prefix maximum
    } else {
      const u32 lim = n - max(i, j);  // bytes available in the shorter suffix
      u32 l = 0;
      bool open = true;
      while (open && l < lim && l < SHORT_BYTES) {
        const u32 nb = min(8u, lim - l);
        const u32 e = equal_bytes(load8(T + i + l, nb), load8(T + j + l, nb), nb);
        l += e;
        open = e == 8u;
      }
      if (open && l < lim) {
        out = REACH_LONG;
        aux[i] = j;
        atomicAdd(n_long, 1u);
      } else {
        out = i + l;
      }
    }
    a[i] = out;
  }
}

// One warp per 32 positions; the unfinished comparisons among them are completed one after the
// other by the whole warp, each lane comparing 4 x 16 bytes per step.
__global__ void __launch_bounds__(256) k_long(const u8 *__restrict__ T, u32 n, u32 *__restrict__ a,
                                              const u32 *__restrict__ aux) {
  const u32 lane = threadIdx.x & 31u;
  const u32 warps = (gridDim.x * blockDim.x) >> 5;
  for (u32 base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < n; base += warps * 32u) {
    const u32 i_me = base + lane;
    u32 todo = __ballot_sync(0xffffffffu, i_me < n && a[i_me] == REACH_LONG);
    while (todo) {
      const u32 src = (u32)__ffs(todo) - 1u;
      todo &= todo - 1u;
      const u32 i = base + src;
      const u32 j = aux[i];
      const u32 lim = n - max(i, j);
      u32 l = SHORT_BYTES;  // known equal so far
      for (;;) {
        // lane covers bytes [l + 64*lane, l + 64*lane + 64)
        const u32 off = l + 64u * lane;
        u32 eq = 64u;  // equal bytes in my stretch (64 = all)
        if (off >= lim) {
          eq = 0u;
        } else {
          const u32 len = min(64u, lim - off);
          u32 k = 0;
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            const u32 o = 8u * (u32)s;
            if (k == o && o < len) {
              const u32 nb = min(8u, len - o);
              k += equal_bytes(load8(T + i + off + o, nb), load8(T + j + off + o, nb), nb);
            }
          }
          eq = k;
        }
        const u32 stop = __ballot_sync(0xffffffffu, eq < 64u);
        if (stop) {
          const u32 first = (u32)__ffs(stop) - 1u;
          const u32 e = __shfl_sync(0xffffffffu, eq, first);
          l += 64u * first + e;
          break;
        }
        l += 64u * 32u;
      }
      if (lane == 0) a[i] = i + min(l, lim);
    }
  }
}

// ---- inclusive prefix maximum over a[] (reach values), then PLCP[i] = max - i, in place -------
__global__ void __launch_bounds__(SCAN_THREADS) k_reach_tiles(const u32 *__restrict__ a, u32 n, u32 *__restrict__ tile_max) {
  __shared__ u32 s_w[SCAN_THREADS / 32];
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 l0 = blockIdx.x * (u32)SCAN_TILE + tid * SCAN_IPT;
  u32 m = 0;
  if (l0 + SCAN_IPT <= n) {
#pragma unroll
    for (int j = 0; j < SCAN_IPT; j += 4) {
      const uint4 v = *reinterpret_cast<const uint4 *>(a + l0 + j);
      m = max(max(m, v.x), max(max(v.y, v.z), v.w));
    }
  } else {
    for (u32 j = 0; j < (u32)SCAN_IPT; ++j)
      if (l0 + j < n) m = max(m, a[l0 + j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_w[warp] = m;
  __syncthreads();
  if (tid == 0) {
    u32 t = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t = max(t, s_w[w]);
    tile_max[blockIdx.x] = t;
  }
}

// tile_carry[t] = max over tiles t' < t of tile_max[t']  (one block)
__global__ void __launch_bounds__(1024) k_reach_spine(const u32 *__restrict__ tile_max, u32 *__restrict__ tile_carry, u32 tiles) {
  __shared__ u32 s_v[1024];
  const u32 t = threadIdx.x;
  const u32 per = (tiles + 1023u) / 1024u;
  const u32 lo = min(tiles, t * per), hi = min(tiles, lo + per);
  u32 m = 0;
  for (u32 i = lo; i < hi; ++i) m = max(m, tile_max[i]);
  s_v[t] = m;
  __syncthreads();
  u32 x = (t > 0) ? s_v[t - 1] : 0u;
  __syncthreads();
  s_v[t] = x;
  __syncthreads();
  for (u32 o = 1; o < 1024; o <<= 1) {
    const u32 y = (t >= o) ? s_v[t - o] : 0u;
    __syncthreads();
    s_v[t] = max(s_v[t], y);
    __syncthreads();
  }
  u32 carry = s_v[t];
  for (u32 i = lo; i < hi; ++i) {
    tile_carry[i] = carry;
    carry = max(carry, tile_max[i]);
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_reach_apply(u32 *__restrict__ a, u32 n, const u32 *__restrict__ tile_carry) {
  __shared__ u32 s_w[SCAN_THREADS / 32];
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 l0 = blockIdx.x * (u32)SCAN_TILE + tid * SCAN_IPT;
  u32 v[SCAN_IPT];
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) v[j] = (l0 + j < n) ? a[l0 + j] : 0u;
  u32 m = 0;
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) m = max(m, v[j]);
  u32 inc = m;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, y);
  }
  if (lane == 31) s_w[warp] = inc;
  u32 run = __shfl_up_sync(0xffffffffu, inc, 1);  // max over lower lanes
  if (lane == 0) run = 0;
  __syncthreads();
  for (u32 w = 0; w < warp; ++w) run = max(run, s_w[w]);
  run = max(run, tile_carry[blockIdx.x]);
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) {
    run = max(run, v[j]);
    if (l0 + j < n) a[l0 + j] = run - (l0 + j);  // PLCP
  }
}

__global__ void __launch_bounds__(256) k_lcp_gather(const i32 *__restrict__ SA, const u32 *__restrict__ plcp, u32 n,
                                                    i32 *__restrict__ LCP) {
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) LCP[j] = (i32)__ldg(plcp + (u32)SA[j]);
}

}  // namespace

namespace {
struct LcpLayout { u32 *a, *tile_max, *tile_carry, *n_long; u32 tiles; size_t total; };
LcpLayout lcp_layout(char *base, u32 n) {
  LcpLayout y;
  Carve c{base, 0};
  y.tiles = (u32)div_up((size_t)n, (size_t)SCAN_TILE);
  y.a = c.take<u32>(n);
  y.tile_max = c.take<u32>(y.tiles + 1);
  y.tile_carry = c.take<u32>(y.tiles + 1);
  y.n_long = c.take<u32>(64);
  y.total = c.used;
  return y;
}
}  // namespace

size_t lcp_workspace_bytes(u32 n) { return lcp_layout(nullptr, n == 0 ? 1 : n).total + 256; }

int lcp_device(const u8 *d_T, const i32 *d_SA, u32 n, i32 *d_LCP, void *workspace, size_t workspace_bytes,
               cudaStream_t st) {
  if (n == 0) return GSA_OK;
  char *owned = nullptr;
  const size_t need = lcp_workspace_bytes(n);
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
  const LcpLayout y = lcp_layout(static_cast<char *>(workspace) + mis, n);
  u32 *a = y.a, *tile_max = y.tile_max, *tile_carry = y.tile_carry, *n_long = y.n_long;
  const u32 tiles = y.tiles;
  u32 *aux = reinterpret_cast<u32 *>(d_LCP);  // the output buffer is free until the final gather

  int dev = 0, sms = kDefaultSMs;
  GSA_TRY(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const u32 blocks = (u32)std::min<u64>((u64)sms * 8, div_up(n, 256));
  GSA_TRY(cudaMemsetAsync(n_long, 0, sizeof(u32), st));
  k_phi<<<blocks, 256, 0, st>>>(d_SA, n, a);
  GSA_TRY(cudaGetLastError());
  k_irreducible<<<blocks, 256, 0, st>>>(d_T, n, a, aux, n_long);
  GSA_TRY(cudaGetLastError());
  u32 nl = 0;
  GSA_TRY(cudaMemcpyAsync(&nl, n_long, sizeof(u32), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaStreamSynchronize(st));
  if (nl) {
    k_long<<<blocks, 256, 0, st>>>(d_T, n, a, aux);
    GSA_TRY(cudaGetLastError());
  }
  k_reach_tiles<<<tiles, SCAN_THREADS, 0, st>>>(a, n, tile_max);
  GSA_TRY(cudaGetLastError());
  k_reach_spine<<<1, 1024, 0, st>>>(tile_max, tile_carry, tiles);
  GSA_TRY(cudaGetLastError());
  k_reach_apply<<<tiles, SCAN_THREADS, 0, st>>>(a, n, tile_carry);
  GSA_TRY(cudaGetLastError());
  k_lcp_gather<<<blocks, 256, 0, st>>>(d_SA, a, n, d_LCP);
  GSA_TRY(cudaGetLastError());
  GSA_TRY(cudaStreamSynchronize(st));
  return GSA_OK;
}

}  // namespace gsa
