// sa_build.cu -- suffix-array construction on the GPU: prefix doubling with
// singleton-group discard.  Replaces the *result* of divsufsort()
// (reference: crates/cdivsufsort/c-sources/divsufsort.c:331-370,
// crates/divsufsort/src/divsufsort.rs:3-37); none of the induced-sorting code is
// reproduced.
//
// Round 0   text bytes -> dense codes (b bits) -> bit-packed stream; key(i) = the next
//           k = floor(64/b) symbols of suffix i, zero padded; LSD radix sort of
//           (key, i); adjacent-key-differs flags + scan give group heads;
//           rank[i] = head slot + 1 (0 is the end-of-text sentinel).
// Round r   live (not yet unique) suffixes only: key = (group ordinal, rank[i + h]);
//           sort, re-flag, re-rank, finalise singletons into SA, compact the rest.
//
// Data layout in HBM (n = text length, L = live suffixes, all arrays contiguous):
//   packed  u64[n*b/64 + 2]   keys u64[n] x2   vals(suffix) u32[n] x2
//   pos u32[n] x2 (SA slot of each live element)   ord u32[n] (group ordinal)
//   rank u32[n]   SA i32[n] (caller's)   + histogram / look-back status scratch.
#include "builder.h"
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "radix.cuh"

namespace gsa {

// ------------------------------------------------------------------------------------
// Alphabet scan: which byte values occur.  One shared-memory store per byte, no atomics.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_byte_presence(const u8 *__restrict__ T, u32 n, u32 *__restrict__ present) {
  __shared__ u8 sp[256];
  sp[threadIdx.x] = 0;
  __syncthreads();
  const u32 nvec = n / 16u;
  const uint4 *T4 = reinterpret_cast<const uint4 *>(T);  // cudaMalloc'd / 256 B aligned text
  for (u32 v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
    const uint4 x = ld_stream_u128(T4 + v);
    const u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sp[w[q] & 255u] = 1;
      sp[(w[q] >> 8) & 255u] = 1;
      sp[(w[q] >> 16) & 255u] = 1;
      sp[w[q] >> 24] = 1;
    }
  }
  if (blockIdx.x == 0)
    for (u32 i = nvec * 16u + threadIdx.x; i < n; i += blockDim.x) sp[T[i]] = 1;
  __syncthreads();
  if (sp[threadIdx.x]) present[threadIdx.x] = 1u;
}

// ------------------------------------------------------------------------------------
// Bit-pack the text: symbol i = code[T[i]] on b bits, big-endian bit order.
// One thread per output word; words past the text end are written as zero.
// ------------------------------------------------------------------------------------
struct CodeMap { u8 code[256]; };

__global__ void __launch_bounds__(256) k_pack(const u8 *__restrict__ T, u32 n, u32 b, const CodeMap cm,
                                              u64 *__restrict__ packed, u64 nwords) {
  __shared__ u8 scode[256];
  scode[threadIdx.x] = cm.code[threadIdx.x];
  __syncthreads();
  const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nwords) return;
  const u64 bit0 = q * 64u;
  u64 i = bit0 / b;  // first symbol touching this word (it may start in the previous word)
  u64 word = 0;
  for (; i < n; ++i) {
    const i64 off = (i64)(i * b) - (i64)bit0;  // bit offset of the symbol from the word's MSB
    if (off >= 64) break;
    const u64 c = scode[T[i]];
    const int sh = 64 - (int)b - (int)off;     // in (-(b), 64)
    word |= (sh >= 0) ? (c << sh) : (c >> (-sh));
  }
  packed[q] = word;
}

// ------------------------------------------------------------------------------------
// Histograms.
// ------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_hist0(const KeyGen g, int npass, u32 *__restrict__ ghist) {
  __shared__ u32 shist[MAX_PASSES * RADIX];
  for (int i = threadIdx.x; i < npass * RADIX; i += THREADS) shist[i] = 0;
  __syncthreads();
  const u32 stride = gridDim.x * THREADS;
  const u32 iters = (g.n + stride - 1) / stride;  // same trip count for every lane
  u32 j = blockIdx.x * THREADS + threadIdx.x;
  for (u32 it = 0; it < iters; ++it, j += stride) {
    const bool valid = j < g.n;
    u64 key = 0;
    u32 s;
    if (valid) gen_key0(g, j, key, s);
    hist_add(shist, key, valid, npass);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * RADIX; i += THREADS)
    if (shist[i]) atomicAdd(&ghist[i], shist[i]);
}

// Rounds >= 1: build the sort key of every live suffix and histogram its digits.
//   key = (ordinal of the suffix's group among live groups) << rank_bits | rank[i + h]
//   (rank 0 = "i + h is past the end", which sorts first: a proper prefix is smaller).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_gather_hist(const u32 *__restrict__ sufx, const u32 *__restrict__ ord,
                                                         const u32 *__restrict__ rank, u64 *__restrict__ keys, u32 L,
                                                         u32 n, u64 h, u32 rank_bits, int npass,
                                                         u32 *__restrict__ ghist) {
  __shared__ u32 shist[MAX_PASSES * RADIX];
  for (int i = threadIdx.x; i < npass * RADIX; i += THREADS) shist[i] = 0;
  __syncthreads();
  const u32 stride = gridDim.x * THREADS;
  const u32 iters = (L + stride - 1) / stride;
  u32 l = blockIdx.x * THREADS + threadIdx.x;
  for (u32 it = 0; it < iters; ++it, l += stride) {
    const bool valid = l < L;
    u64 key = 0;
    if (valid) {
      const u64 t = (u64)ld_stream_u32(sufx + l) + h;
      const u32 r2 = (t < n) ? __ldg(rank + t) : 0u;
      key = ((u64)ld_stream_u32(ord + l) << rank_bits) | r2;
      keys[l] = key;
    }
    hist_add(shist, key, valid, npass);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * RADIX; i += THREADS)
    if (shist[i]) atomicAdd(&ghist[i], shist[i]);
}

// ------------------------------------------------------------------------------------
// Rank rebuild + singleton finalisation + live-set compaction, one pass, chained scan.
//   flag[l]  = key[l] != key[l-1]            (round 0: also around short suffixes)
//   head[l]  = SA slot of the nearest flagged element at or before l   (max-scan)
//   rank[suffix[l]] = head[l] + 1
//   singleton (flag[l] && flag[l+1])  -> SA[pos[l]] = suffix[l], dropped from the live set
//   otherwise -> appended to the next live set with its slot and its group's ordinal
// Tile prefix (head, survivors, surviving groups) travels through a decoupled look-back
// over two self-validating 64-bit status words per tile.
// ------------------------------------------------------------------------------------
struct RebuildArgs {
  const u64 *keys;
  const u32 *sufx;
  const u32 *pos_in;   // null in round 0 (slot == index)
  u32 L;
  u32 short_from;      // round 0: suffix indices >= short_from are short (forced singletons)
  u32 *rank;
  i32 *SA;
  u32 *pos_out;
  u32 *sufx_out;
  u32 *ord_out;
  ulonglong2 *status;  // [tiles] x = flag(2) | head+1 ; y = flag(2) | survivors(31) | groups(31)
  RoundResult *result;
};

constexpr u64 ST_AGG = 1ull << 62;
constexpr u64 ST_PRE = 2ull << 62;
constexpr u64 ST_FLAG = 3ull << 62;

// One 16-byte, 16-byte-aligned access per tile descriptor (a single memory transaction on
// the hardware).  Both halves carry the state flag, so a torn read -- should one ever
// happen -- is seen as "flags differ" and simply retried.
__device__ __forceinline__ ulonglong2 ld_status(const ulonglong2 *p) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(ulonglong2 *p, u64 x, u64 y) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}

template <int THREADS, int IPT, bool ROUND0>
__global__ void __launch_bounds__(THREADS) k_rebuild(const RebuildArgs a) {
  constexpr int WARPS = THREADS / 32;
  constexpr int TILE = THREADS * IPT;
  __shared__ u32 s_wh[WARPS], s_wc[WARPS], s_wg[WARPS];
  __shared__ u32 s_pre[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Tiles are taken in blockIdx order: the look-back only ever waits on lower-numbered
  // blocks, which the hardware dispatches first.
  const u32 tile = blockIdx.x;
  const u32 L = a.L;
  const u32 l0 = tile * (u32)TILE + (u32)tid * IPT;

  // ---- load this thread's IPT consecutive elements (+ one neighbour on each side) -------
  u64 kx[IPT + 2];  // kx[j+1] = key of element l0+j
  u32 sx[IPT + 2];
  u32 px[IPT];
  if (l0 + IPT <= L) {
    static_assert(IPT % 4 == 0, "vector loads");
#pragma unroll
    for (int j = 0; j < IPT; j += 2) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(a.keys + l0 + j);
      kx[j + 1] = v.x;
      kx[j + 2] = v.y;
    }
#pragma unroll
    for (int j = 0; j < IPT; j += 4) {
      const uint4 v = *reinterpret_cast<const uint4 *>(a.sufx + l0 + j);
      sx[j + 1] = v.x; sx[j + 2] = v.y; sx[j + 3] = v.z; sx[j + 4] = v.w;
    }
    if (!ROUND0) {
#pragma unroll
      for (int j = 0; j < IPT; j += 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(a.pos_in + l0 + j);
        px[j] = v.x; px[j + 1] = v.y; px[j + 2] = v.z; px[j + 3] = v.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const bool ok = l0 + j < L;
      kx[j + 1] = ok ? a.keys[l0 + j] : 0;
      sx[j + 1] = ok ? a.sufx[l0 + j] : 0;
      if (!ROUND0) px[j] = ok ? a.pos_in[l0 + j] : 0;
    }
  }
  if (ROUND0) {
#pragma unroll
    for (int j = 0; j < IPT; ++j) px[j] = l0 + j;
  }
  const bool has_prev = (l0 > 0) && (l0 <= L);
  const bool has_next = (l0 + IPT < L);
  kx[0] = has_prev ? a.keys[l0 - 1] : 0;
  sx[0] = has_prev ? a.sufx[l0 - 1] : 0;
  kx[IPT + 1] = has_next ? a.keys[l0 + IPT] : 0;
  sx[IPT + 1] = has_next ? a.sufx[l0 + IPT] : 0;

  // ---- flags f[j] for elements l0+j, j = 0..IPT (bit j); beyond-the-end counts as flagged ----
  u32 f = 0;
#pragma unroll
  for (int j = 0; j <= IPT; ++j) {
    const u32 l = l0 + j;
    bool fl = (l == 0) || (l >= L) || (kx[j + 1] != kx[j]);
    if (ROUND0) fl = fl || (sx[j + 1] >= a.short_from) || (sx[j] >= a.short_from);
    f |= (fl ? 1u : 0u) << j;
  }
  u32 nvalid = 0;
  if (l0 < L) nvalid = min((u32)IPT, L - l0);

  // ---- thread aggregates ------------------------------------------------------------------
  u32 th = 0, tc = 0, tg = 0;  // last flagged slot + 1, survivors, surviving group heads
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    if ((u32)j < nvalid) {
      const bool fj = (f >> j) & 1u, fn = (f >> (j + 1)) & 1u;
      if (fj) th = px[j] + 1u;
      if (!(fj && fn)) ++tc;
      if (fj && !fn) ++tg;
    }
  }
  // ---- block exclusive scan of (max, sum, sum) -------------------------------------------------
  u32 ih = th, ic = tc, ig = tg;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 yh = __shfl_up_sync(0xffffffffu, ih, o);
    const u32 yc = __shfl_up_sync(0xffffffffu, ic, o);
    const u32 yg = __shfl_up_sync(0xffffffffu, ig, o);
    if (lane >= o) { ih = max(ih, yh); ic += yc; ig += yg; }
  }
  if (lane == 31) { s_wh[warp] = ih; s_wc[warp] = ic; s_wg[warp] = ig; }
  // exclusive within warp
  u32 eh = __shfl_up_sync(0xffffffffu, ih, 1), ec = ic - tc, eg = ig - tg;
  if (lane == 0) eh = 0;
  __syncthreads();
  u32 bh = 0, bc = 0, bg = 0;  // block totals (all warps)
  {
    u32 ph = 0, pc = 0, pg = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      if (w == warp) { ph = bh; pc = bc; pg = bg; }
      bh = max(bh, s_wh[w]); bc += s_wc[w]; bg += s_wg[w];
    }
    eh = max(eh, ph); ec += pc; eg += pg;
  }

  // ---- tile prefix by decoupled look-back (warp 0, 32 predecessors per round trip) ----------
  if (warp == 0) {
    u32 xh = 0, xc = 0, xg = 0;  // exclusive prefix over preceding tiles
    if (tile == 0) {
      if (lane == 0) st_status(a.status, ST_PRE | bh, ST_PRE | ((u64)bc << 31) | bg);
    } else {
      if (lane == 0) st_status(a.status + tile, ST_AGG | bh, ST_AGG | ((u64)bc << 31) | bg);
      i64 look = (i64)tile - 1 - lane;  // lane 0 inspects the nearest predecessor
      for (;;) {
        ulonglong2 sv = make_ulonglong2(ST_PRE, ST_PRE);  // virtual tiles before tile 0: identity prefix
        if (look >= 0) {
          do { sv = ld_status(a.status + look); } while ((sv.x & ST_FLAG) == 0 || (sv.x & ST_FLAG) != (sv.y & ST_FLAG));
        }
        const u32 pre = __ballot_sync(0xffffffffu, (sv.x & ST_FLAG) == ST_PRE);
        const int first = pre ? (__ffs(pre) - 1) : 32;  // nearest tile holding an inclusive prefix
        u32 vh = 0, vc = 0, vg = 0;
        if (lane <= first) {
          vh = (u32)(sv.x & 0xffffffffull);
          vc = (u32)((sv.y >> 31) & 0x7fffffffull);
          vg = (u32)(sv.y & 0x7fffffffull);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          vh = max(vh, __shfl_xor_sync(0xffffffffu, vh, o));
          vc += __shfl_xor_sync(0xffffffffu, vc, o);
          vg += __shfl_xor_sync(0xffffffffu, vg, o);
        }
        xh = max(xh, vh); xc += vc; xg += vg;
        if (pre) break;
        look -= 32;
      }
      if (lane == 0) st_status(a.status + tile, ST_PRE | max(xh, bh), ST_PRE | ((u64)(xc + bc) << 31) | (xg + bg));
    }
    if (lane == 0) {
      s_pre[0] = xh; s_pre[1] = xc; s_pre[2] = xg;
      if ((u64)(tile + 1) * TILE >= L) {  // last tile: totals of the round
        a.result->live_out = xc + bc;
        a.result->groups_out = xg + bg;
      }
    }
  }
  __syncthreads();

  // ---- emit ------------------------------------------------------------------------------------
  u32 head = max(s_pre[0], eh);  // head slot + 1
  u32 c = s_pre[1] + ec;
  u32 g = s_pre[2] + eg;
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    if ((u32)j < nvalid) {
      const bool fj = (f >> j) & 1u, fn = (f >> (j + 1)) & 1u;
      if (fj) head = px[j] + 1u;
      a.rank[sx[j + 1]] = head;
      if (fj && fn) {
        a.SA[px[j]] = (i32)sx[j + 1];
      } else {
        if (fj) ++g;
        a.pos_out[c] = px[j];
        a.sufx_out[c] = sx[j + 1];
        a.ord_out[c] = g - 1u;
        ++c;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Host driver
// ------------------------------------------------------------------------------------
namespace {

constexpr int PASS_THREADS = 256;
constexpr int PASS_IPT = 16;
constexpr int PASS_TILE = PASS_THREADS * PASS_IPT;
constexpr int RB_THREADS = 512;
constexpr int RB_IPT = 8;
constexpr int RB_TILE = RB_THREADS * RB_IPT;
constexpr int HIST_THREADS = 512;

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

struct Layout {
  u64 *packed; u64 packed_words;
  u64 *keys[2]; u32 *vals[2]; u32 *pos[2]; u32 *ord; u32 *rank;
  u32 *ghist;      // [MAX_PASSES][256]
  u32 *bin_base;   // [MAX_PASSES][256]
  u32 *present;    // [256]
  u32 *skip_mask;  // [1]
  RoundResult *result;
  u32 *pass_status; size_t pass_status_words;  // counter at word 0 (256-word header), then [tiles][256]
  ulonglong2 *rb_status; size_t rb_status_words;  // one 16-byte descriptor per rebuild tile
  size_t total;
};

Layout make_layout(char *base, u32 n) {
  Layout y;
  Carve c{base, 0};
  const size_t N = n;
  y.packed_words = N / 8 + 4;  // b <= 8 bits per symbol
  y.packed = c.take<u64>(y.packed_words);
  y.keys[0] = c.take<u64>(N); y.keys[1] = c.take<u64>(N);
  y.vals[0] = c.take<u32>(N); y.vals[1] = c.take<u32>(N);
  y.pos[0] = c.take<u32>(N);  y.pos[1] = c.take<u32>(N);
  y.ord = c.take<u32>(N);
  y.rank = c.take<u32>(N);
  y.ghist = c.take<u32>(MAX_PASSES * RADIX);
  y.bin_base = c.take<u32>(MAX_PASSES * RADIX);
  y.present = c.take<u32>(256);
  y.skip_mask = c.take<u32>(64);
  y.result = c.take<RoundResult>(16);
  const size_t ptiles = div_up(N, PASS_TILE);
  y.pass_status_words = 256 + ptiles * RADIX;
  y.pass_status = c.take<u32>(y.pass_status_words);
  const size_t rtiles = div_up(N, RB_TILE);
  y.rb_status_words = rtiles + 1;
  y.rb_status = c.take<ulonglong2>(y.rb_status_words);
  y.total = c.used;
  return y;
}

}  // namespace

size_t build_workspace_bytes(u32 n) { return make_layout(nullptr, n == 0 ? 1 : n).total + 256; }

#define KLAUNCH_CHECK() GSA_TRY(cudaGetLastError())

// Event pairs around every k_radix_pass launch: the roofline of the dominant kernel is
// reported from these (bytes moved per launch / launch duration), not from round totals.
struct PassTimer {
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  ~PassTimer() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
  int next(cudaEvent_t *out) {
    if (used == ev.size()) {
      cudaEvent_t e;
      GSA_TRY(cudaEventCreate(&e));
      ev.push_back(e);
    }
    *out = ev[used++];
    return GSA_OK;
  }
  // call after the stream has been synchronised
  float drain() {
    float total = 0.f;
    for (size_t i = 0; i + 1 < used; i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) total += ms;
    }
    used = 0;
    return total;
  }
};

// Runs the radix passes for digits [0, npass) on `L` elements whose histograms are already
// in y.ghist.  `cur` is the buffer index holding the input (ignored when gen != null: the
// first pass then generates the keys and writes buffer 0).  Returns the index of the buffer
// holding the sorted pairs.
static int run_passes(const Layout &y, u32 L, int npass, int cur, const KeyGen *gen, cudaStream_t st,
                      gsa_build_stats *stats, PassTimer &timer, int *cur_out, u32 *passes_done) {
  GSA_TRY(cudaMemsetAsync(y.skip_mask, 0, sizeof(u32), st));
  k_scan_hist<<<npass, RADIX, 0, st>>>(y.ghist, y.bin_base, L, y.skip_mask);
  KLAUNCH_CHECK();
  u32 skip = 0;
  GSA_TRY(cudaMemcpyAsync(&skip, y.skip_mask, sizeof(u32), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaStreamSynchronize(st));
  const u32 tiles = (u32)div_up(L, PASS_TILE);
  const size_t smem = PassCfg<PASS_THREADS, PASS_IPT>::SMEM;
  bool need_gen = gen != nullptr;
  u32 done = 0;
  for (int p = 0; p < npass; ++p) {
    if (((skip >> p) & 1u) && !(need_gen && p == npass - 1)) continue;  // constant digit: identity pass
    GSA_TRY(cudaMemsetAsync(y.pass_status, 0, (256 + (size_t)tiles * RADIX) * sizeof(u32), st));
    PassArgs a;
    a.n = L;
    a.shift = 8u * (u32)p;
    a.bin_base = y.bin_base + p * RADIX;
    a.counter = y.pass_status;
    a.status = y.pass_status + 256;
    cudaEvent_t t0, t1;
    GSA_TRY_RC(timer.next(&t0));
    GSA_TRY_RC(timer.next(&t1));
    GSA_TRY(cudaEventRecord(t0, st));
    if (need_gen) {
      a.keys_in = nullptr; a.vals_in = nullptr;
      a.keys_out = y.keys[0]; a.vals_out = y.vals[0];
      a.gen = *gen;
      k_radix_pass<PASS_THREADS, PASS_IPT, true><<<tiles, PASS_THREADS, smem, st>>>(a);
      cur = 0;
      need_gen = false;
    } else {
      a.keys_in = y.keys[cur]; a.vals_in = y.vals[cur];
      a.keys_out = y.keys[cur ^ 1]; a.vals_out = y.vals[cur ^ 1];
      a.gen = KeyGen{};
      k_radix_pass<PASS_THREADS, PASS_IPT, false><<<tiles, PASS_THREADS, smem, st>>>(a);
      cur ^= 1;
    }
    KLAUNCH_CHECK();
    GSA_TRY(cudaEventRecord(t1, st));
    ++done;
    if (stats) {
      stats->radix_pass_launches++;
      stats->radix_pass_elements += L;
      stats->kernel_launches++;
    }
  }
  if (stats) stats->kernel_launches++;  // k_scan_hist
  *cur_out = cur;
  *passes_done = done;
  return GSA_OK;
}

int build_sa_device(const u8 *d_T, i32 *d_SA, u32 n, void *workspace, size_t workspace_bytes, cudaStream_t st,
                    gsa_build_stats *stats) {
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n == 0) return GSA_OK;
  static_assert(PASS_THREADS >= RADIX, "");
  {
    // opt in to > 48 KB dynamic shared memory (idempotent, per device)
    const int smem = (int)PassCfg<PASS_THREADS, PASS_IPT>::SMEM;
    GSA_TRY(cudaFuncSetAttribute(k_radix_pass<PASS_THREADS, PASS_IPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    GSA_TRY(cudaFuncSetAttribute(k_radix_pass<PASS_THREADS, PASS_IPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  char *owned = nullptr;
  const size_t need = build_workspace_bytes(n);
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
  struct Free { char *p; ~Free() { if (p) cudaFree(p); } } free_guard{owned};
  const size_t mis = (256 - (reinterpret_cast<uintptr_t>(workspace) & 255)) & 255;
  const Layout y = make_layout(static_cast<char *>(workspace) + mis, n);

  int sms = kDefaultSMs;
  {
    int dev = 0;
    GSA_TRY(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // The rank gather / scatter of the doubling rounds touches 4 bytes per 32-byte sector at
  // random; ask L2 not to widen those misses to 64/128-byte DRAM fetches (ncu: 126 B of DRAM
  // read per gathered rank at the default setting).  Restored before returning.
  struct L2Fetch {
    size_t old = 0;
    bool changed = false;
    L2Fetch() {
      const char *e = getenv("GSA_L2_FETCH");
      const size_t want = e ? (size_t)atoi(e) : 32;
      if (want == 0) return;  // GSA_L2_FETCH=0: leave the device setting alone
      if (cudaDeviceGetLimit(&old, cudaLimitMaxL2FetchGranularity) != cudaSuccess) { cudaGetLastError(); return; }
      if (old != want && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, want) == cudaSuccess) changed = true;
      else cudaGetLastError();
    }
    ~L2Fetch() { if (changed) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, old); cudaGetLastError(); } }
  } l2fetch_guard;
  cudaEvent_t ev[4];
  for (auto &e : ev) GSA_TRY(cudaEventCreate(&e));
  struct EvFree { cudaEvent_t *e; ~EvFree() { for (int i = 0; i < 4; ++i) cudaEventDestroy(e[i]); } } ev_guard{ev};
  cudaEvent_t ev_all0, ev_all1;
  GSA_TRY(cudaEventCreate(&ev_all0));
  GSA_TRY(cudaEventCreate(&ev_all1));
  struct Ev2 { cudaEvent_t a, b; ~Ev2() { cudaEventDestroy(a); cudaEventDestroy(b); } } ev2_guard{ev_all0, ev_all1};
  GSA_TRY(cudaEventRecord(ev_all0, st));
  GSA_TRY(cudaEventRecord(ev[0], st));

  // ---- alphabet ---------------------------------------------------------------------------
  GSA_TRY(cudaMemsetAsync(y.present, 0, 256 * sizeof(u32), st));
  {
    const u32 blocks = (u32)std::min<u64>((u64)sms * 8, std::max<u64>(1, div_up(n, 256 * 16)));
    k_byte_presence<<<blocks, 256, 0, st>>>(d_T, n, y.present);
    KLAUNCH_CHECK();
  }
  u32 present[256];
  GSA_TRY(cudaMemcpyAsync(present, y.present, sizeof(present), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaStreamSynchronize(st));
  CodeMap cm;
  u32 sigma = 0;
  for (int c = 0; c < 256; ++c) {
    cm.code[c] = (u8)sigma;
    if (present[c]) ++sigma;
  }
  const u32 b = bits_for(sigma > 1 ? sigma - 1 : 1);  // codes 0..sigma-1
  const u32 k = 64 / b;                               // symbols per round-0 key
  const u32 key_bits = k * b;
  const u32 ns = (k - 1 < n) ? (k - 1) : n;           // short suffixes
  const u64 nwords = ((u64)n * b + 63) / 64 + 2;
  {
    k_pack<<<(u32)div_up(nwords, 256), 256, 0, st>>>(d_T, n, b, cm, y.packed, nwords);
    KLAUNCH_CHECK();
  }
  if (stats) {
    stats->sigma = sigma; stats->bits_per_symbol = b; stats->symbols_per_key = k;
    stats->kernel_launches += 2;
  }

  // ---- round 0 ------------------------------------------------------------------------------
  KeyGen gen{y.packed, n, ns, b, key_bits};
  const int npass0 = (int)div_up(key_bits, 8);
  const u32 hist_blocks = (u32)std::min<u64>((u64)sms * 4, std::max<u64>(1, div_up(n, HIST_THREADS)));
  GSA_TRY(cudaMemsetAsync(y.ghist, 0, MAX_PASSES * RADIX * sizeof(u32), st));
  k_hist0<HIST_THREADS><<<hist_blocks, HIST_THREADS, 0, st>>>(gen, npass0, y.ghist);
  KLAUNCH_CHECK();
  if (stats) stats->kernel_launches++;
  int cur = 0;
  u32 passes = 0;
  PassTimer timer;
  GSA_TRY(cudaEventRecord(ev[1], st));
  GSA_TRY_RC(run_passes(y, n, npass0, 0, &gen, st, stats, timer, &cur, &passes));
  GSA_TRY(cudaEventRecord(ev[2], st));

  auto launch_rebuild = [&](bool round0, u32 L, int kv, int pin, int pout) -> int {
    const u32 tiles = (u32)div_up(L, RB_TILE);
    GSA_TRY(cudaMemsetAsync(y.rb_status, 0, (size_t)tiles * sizeof(ulonglong2), st));
    RebuildArgs r;
    r.keys = y.keys[kv]; r.sufx = y.vals[kv];
    r.pos_in = round0 ? nullptr : y.pos[pin];
    r.L = L;
    r.short_from = round0 ? (n - ns) : 0xffffffffu;
    r.rank = y.rank; r.SA = d_SA;
    r.pos_out = y.pos[pout]; r.sufx_out = y.vals[kv ^ 1]; r.ord_out = y.ord;
    r.status = y.rb_status;
    r.result = y.result;
    if (round0)
      k_rebuild<RB_THREADS, RB_IPT, true><<<tiles, RB_THREADS, 0, st>>>(r);
    else
      k_rebuild<RB_THREADS, RB_IPT, false><<<tiles, RB_THREADS, 0, st>>>(r);
    KLAUNCH_CHECK();
    if (stats) stats->kernel_launches++;
    return GSA_OK;
  };

  RoundResult rr{};
  GSA_TRY_RC(launch_rebuild(true, n, cur, 0, 0));
  GSA_TRY(cudaMemcpyAsync(&rr, y.result, sizeof(rr), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaEventRecord(ev[3], st));
  GSA_TRY(cudaStreamSynchronize(st));
  u32 round = 0;
  auto log_round = [&](u64 depth, u64 live, u32 groups, u32 kb, u32 np) {
    if (!stats || round >= GSA_MAX_ROUNDS) return;
    gsa_round_stat &s = stats->round[round];
    s.depth = depth; s.live = live; s.groups = groups; s.key_bits = kb; s.passes = np;
    cudaEventElapsedTime(&s.ms_total, ev[0], ev[3]);
    cudaEventElapsedTime(&s.ms_sort, ev[1], ev[2]);
    stats->ms_radix_passes += timer.drain();
    stats->rounds = round + 1;
  };
  log_round(k, n, 0, key_bits, passes);

  // ---- doubling rounds ------------------------------------------------------------------------
  int vcur = cur ^ 1;  // buffer holding the live suffixes
  int pcur = 0;        // pos buffer holding their slots
  u64 h = k;           // suffixes are sorted by their first h symbols
  const u32 rank_bits = bits_for(n);  // ranks are 0..n
  while (rr.live_out > 0) {
    ++round;
    const u32 L = rr.live_out, G = rr.groups_out;
    const u32 kb = bits_for(G > 0 ? G - 1 : 0) + rank_bits;
    const int npass = (int)div_up(kb, 8);
    GSA_TRY(cudaEventRecord(ev[0], st));
    GSA_TRY(cudaMemsetAsync(y.ghist, 0, MAX_PASSES * RADIX * sizeof(u32), st));
    const u32 gblocks = (u32)std::min<u64>((u64)sms * 4, std::max<u64>(1, div_up(L, HIST_THREADS)));
    k_gather_hist<HIST_THREADS><<<gblocks, HIST_THREADS, 0, st>>>(y.vals[vcur], y.ord, y.rank, y.keys[vcur], L, n, h,
                                                                 rank_bits, npass, y.ghist);
    KLAUNCH_CHECK();
    if (stats) stats->kernel_launches++;
    GSA_TRY(cudaEventRecord(ev[1], st));
    GSA_TRY_RC(run_passes(y, L, npass, vcur, nullptr, st, stats, timer, &cur, &passes));
    GSA_TRY(cudaEventRecord(ev[2], st));
    GSA_TRY_RC(launch_rebuild(false, L, cur, pcur, pcur ^ 1));
    GSA_TRY(cudaMemcpyAsync(&rr, y.result, sizeof(rr), cudaMemcpyDeviceToHost, st));
    GSA_TRY(cudaEventRecord(ev[3], st));
    GSA_TRY(cudaStreamSynchronize(st));
    h *= 2;
    log_round(h, L, G, kb, passes);
    vcur = cur ^ 1;
    pcur ^= 1;
    if (round > 64) { set_error("prefix doubling did not converge", __FILE__, __LINE__); return GSA_ECUDA; }
  }
  GSA_TRY(cudaEventRecord(ev_all1, st));
  GSA_TRY(cudaStreamSynchronize(st));
  if (stats) cudaEventElapsedTime(&stats->ms_total, ev_all0, ev_all1);
  return GSA_OK;
}

}  // namespace gsa
