// radix.cuh -- hand-written LSD radix sort building blocks for (u64 key, u32 value)
// pairs on sm_100a.  No CUB / Thrust.
//
//   hist_add()          warp-aggregated multi-digit histogram update (runs of equal adjacent
//                       keys are added once, by ballot; no match.any)
//   k_scan_hist         per-digit exclusive scan of the global histograms + detection
//                       of constant digits (their pass is skipped by the host)
//   k_radix_pass        one "onesweep" pass: a tile is ranked in shared memory with
//                       warp ballots / match.any, its per-digit counts are chained to the
//                       preceding tiles by a decoupled look-back, and keys/values are
//                       scattered in digit runs.  Optionally generates the round-0 keys
//                       on the fly from the bit-packed text (GEN) so they are never
//                       materialised unsorted.
//
// Traffic per pass and element: 12 B read + 12 B written (8 B key + 4 B value) -- the
// HBM roofline this kernel is measured against ("kernels").
#pragma once
#include "common.cuh"

namespace gsa {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;
#ifndef GSA_PASS_DYNAMIC_TILES
#define GSA_PASS_DYNAMIC_TILES 1
#endif

// ---------------------------------------------------------------------------------
// Round-0 key generation from the bit-packed symbol stream.
//   packed : big-endian bit stream, `b` bits per symbol, symbol i at bits [i*b, i*b+b)
//            counted from the MSB of word 0; zero beyond the text end (>= 2 spare words)
//   element j of the initial sequence is suffix i(j): the `ns` short suffixes (fewer than
//   symbols_per_key symbols left) come first, shortest first, then 0,1,2,...  A stable
//   sort keeps a short suffix in front of every longer suffix that shares its zero-padded
//   key, which is exactly "a proper prefix sorts first".
// ---------------------------------------------------------------------------------
struct KeyGen {
  const u64 *packed;
  u32 n;         // text length
  u32 ns;        // number of short suffixes = min(symbols_per_key - 1, n)
  u32 b;         // bits per symbol
  u32 key_bits;  // symbols_per_key * b  (1..64)
};

#ifdef __CUDACC__
// Round-0 key of suffix i: its first key_bits / b symbols, zero padded past the text end.
__device__ __forceinline__ u64 key_of_suffix(const KeyGen &g, u32 i) {
  const u64 bit = (u64)i * g.b;
  const u64 w = bit >> 6;
  const u32 s = (u32)bit & 63u;
  const u64 w0 = __ldg(g.packed + w);
  const u64 w1 = __ldg(g.packed + w + 1);
  const u64 x = s ? ((w0 << s) | (w1 >> (64u - s))) : w0;
  return x >> (64u - g.key_bits);
}

// Hint: move `bytes` (multiple of 16, 16-byte aligned address) from HBM into L2; no destination, nothing to wait for.
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, u32 bytes) {
  if (bytes != 0u) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void gen_key0(const KeyGen &g, u32 j, u64 &key, u32 &sufx) {
  const u32 i = (j < g.ns) ? (g.n - 1u - j) : (j - g.ns);
  key = key_of_suffix(g, i);
  sufx = i;
}

// Warp-aggregated histogram update for `npass` 8-bit digits of `key` (digit p = bits
// [8p, 8p+8)).  Lanes hold consecutive elements; a run of equal adjacent keys in valid lanes
// (the common case in doubling rounds, where most of a group shares one key) is added once
// by its first lane with the run length, so it does not serialise on one shared-memory
// address.  Distinct neighbours cost a shuffle and two ballots -- no match.any, whose latency
// grows with the number of distinct values in the warp (ncu: 150+ cycles on random keys).
// Must be called by all lanes of the warp.
__device__ __forceinline__ void hist_add(u32 *shist, u64 key, bool valid, int npass) {
  const u32 lane = lane_id();
  const u64 prev = __shfl_up_sync(0xffffffffu, key, 1);
  const u32 vmask = __ballot_sync(0xffffffffu, valid);
  const bool pvalid = lane > 0 && ((vmask >> (lane - 1)) & 1u);
  const bool head = valid && (!pvalid || prev != key);
  const u32 heads = __ballot_sync(0xffffffffu, head);
  if (head) {
    const u32 hi = ~((2u << lane) - 1u);       // lanes above mine
    const u32 stop = (heads | ~vmask) & hi;    // next run head or next empty lane
    const u32 c = (stop ? (u32)(__ffs(stop) - 1) : 32u) - lane;
#pragma unroll 1
    for (int p = 0; p < npass; ++p) atomicAdd(&shist[p * RADIX + (u32)((key >> (8 * p)) & 255u)], c);
  }
}

// ghist[p][256] -> bin_base[p][256] (exclusive scan); sets bit p of *skip_mask when one
// bin of digit p holds all `total` elements (the pass would be the identity).
__global__ void __launch_bounds__(RADIX) k_scan_hist(const u32 *__restrict__ ghist, u32 *__restrict__ bin_base,
                                                     u32 total, u32 *__restrict__ skip_mask,
                                                     const u32 *__restrict__ total_ptr = nullptr) {
  __shared__ u32 wsum[RADIX / 32];
  const int p = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (total_ptr != nullptr) total = *total_ptr;  // element count still on the device (the host has not read it yet)
  const u32 c = ghist[p * RADIX + t];
  if (c == total && total != 0) atomicOr(skip_mask, 1u << p);
  u32 x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  u32 add = 0;
  for (int i = 0; i < w; ++i) add += wsum[i];
  bin_base[p * RADIX + t] = x - c + add;
}

// Tile status word of the pass look-back: 0 = not ready, otherwise value+1 in bits 0..30
// and bit 31 = "inclusive prefix" (else "tile aggregate").  value <= n <= 2^31-2.
__device__ __forceinline__ u32 st_agg(u32 v) { return v + 1u; }
__device__ __forceinline__ u32 st_pre(u32 v) { return (v + 1u) | 0x80000000u; }

struct PassArgs {
  const void *keys_in;  // u64 or u32 keys (template parameter of the pass kernel)
  const u32 *vals_in;
  void *keys_out;
  u32 *vals_out;
  u32 n;               // elements
  u32 shift;           // digit = (key >> shift) & 255
  const u32 *bin_base; // [256] global exclusive offsets of this digit
  u32 *status;         // [tiles][256], zeroed before launch
  u32 *counter;        // dynamic tile id, zeroed before launch
  u32 pf_dist;         // > 0: ask L2 to fetch the pairs of tile (mine + pf_dist) now (cp.async.bulk.prefetch.L2)
  KeyGen gen;          // GEN only
};

template <int THREADS, int IPT, typename KOUT = u64>
struct PassCfg {
  static constexpr int WARPS = THREADS / 32;
  static constexpr int TILE = THREADS * IPT;
  // keys, values, per-warp digit counts (u16: a warp holds at most 32 * IPT elements, a tile offset is below TILE), bin offsets
  static constexpr size_t SMEM = (size_t)TILE * sizeof(KOUT) + (size_t)TILE * 4 + (size_t)WARPS * RADIX * 2 + RADIX * 4;
};

// The lowest lane of every set of equal digits claims `popc(peers)` slots of the warp's bin.  The leaders of one row
// hold DISTINCT digits, hence distinct u16 counters: a plain load + store does what a shared-memory atomic would, and
// a full-warp ATOMS costs ~2 cycles per active lane (64 per row on uniform digits: at 16 rows x 8 warps per tile
// that pipe alone took as long as the tile's share of HBM time), a conflict-free LDS / STS pair a few.  Rows are ordered
// by __syncwarp() (memory ordering among the lanes of the warp).  -DGSA_RANK_ATOMS=1 restores the atomic form.
#ifndef GSA_RANK_ATOMS
#define GSA_RANK_ATOMS 0
#endif
#if GSA_RANK_ATOMS
#define GSA_CLAIM(base, wh, wh32, d, peers, below)                                                                   \
  if ((below) == 0) (base) = (atomicAdd(&(wh32)[(d) >> 1], (u32)__popc(peers) << (16u * ((d) & 1u))) >> (16u * ((d) & 1u))) & 0xffffu
#else
#define GSA_CLAIM(base, wh, wh32, d, peers, below)          \
  do {                                                      \
    if ((below) == 0) {                                     \
      (base) = (wh)[d];                                     \
      (wh)[d] = (u16)((base) + (u32)__popc(peers));         \
    }                                                       \
    __syncwarp();                                           \
  } while (0)
#endif

// Lanes of the warp whose 8-bit digit equals mine, from 8 ballots (cost independent of the
// number of distinct digits in the warp).
__device__ __forceinline__ u32 peers_by_ballot(u32 d) {
  u32 m = 0xffffffffu;
  // Per bit: test -> predicate (LOP3.P), vote, predicate -> 0 / ~0 (SEL), m &= ~(ballot ^ that) (LOP3): 4 instructions.
  // (Written in C -- `bal ^ (bit - 1)` or `p ? bal : ~bal` -- the compiler shifts, masks, compares and decrements: 6.)
#define GSA_PEER_BIT(B)                                                  \
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t, bal;\n\t"                     \
      "and.b32 t, %1, " #B ";\n\t"                                       \
      "setp.ne.u32 p, t, 0;\n\t"                                         \
      "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"                     \
      "selp.b32 t, 0xffffffff, 0, p;\n\t"                                \
      "lop3.b32 %0, %0, bal, t, 0x90;\n\t}"                              \
      : "+r"(m)                                                          \
      : "r"(d))
  GSA_PEER_BIT(1);
  GSA_PEER_BIT(2);
  GSA_PEER_BIT(4);
  GSA_PEER_BIT(8);
  GSA_PEER_BIT(16);
  GSA_PEER_BIT(32);
  GSA_PEER_BIT(64);
  GSA_PEER_BIT(128);
#undef GSA_PEER_BIT
  return m;
}

// KIN / KOUT: width of the keys read and written (u64 everywhere today).  Moving round-0 keys of <= 32 bits as u32
// through all passes but the last (8 + 16 + 16 + 20 instead of 12 + 24 + 24 + 24 bytes per element on rep_1G) was
// measured and changed nothing (round 0: 46.0 -> 46.2 ms, profiles/r2/README.md): the pass is bound by the latency of
// its ranking / look-back chain at 37 % occupancy, not by bytes.
template <int THREADS, int IPT, bool GEN, int MIN_BLOCKS = 3, typename KIN = u64, typename KOUT = u64>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_radix_pass(const PassArgs a) {
  static_assert(THREADS >= RADIX && THREADS % 32 == 0, "one thread per bin is assumed");
  static_assert(IPT % 2 == 0 && THREADS * IPT <= 65535, "ranks, counts and tile offsets are kept as u16");
  constexpr int WARPS = THREADS / 32;
  constexpr int TILE = THREADS * IPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KOUT *skeys = reinterpret_cast<KOUT *>(smem_raw);        // [TILE]
  u32 *svals = reinterpret_cast<u32 *>(skeys + TILE);      // [TILE]
  u16 *whist = reinterpret_cast<u16 *>(svals + TILE);      // [WARPS][256] counts -> tile offsets, two per 32-bit word
  u32 *bin_gofs = reinterpret_cast<u32 *>(whist + WARPS * RADIX);  // [256] global offset - local offset
  __shared__ u32 s_wsum[RADIX / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if GSA_PASS_DYNAMIC_TILES
  __shared__ u32 s_tile;
  if (tid == 0) s_tile = atomicAdd(a.counter, 1u);
#endif
  for (int i = tid; i < WARPS * RADIX / 2; i += THREADS) reinterpret_cast<u32 *>(whist)[i] = 0;
  __syncthreads();
  // Tiles are taken by ticket from a global counter: the look-back then only ever waits on tiles whose CTAs are
  // already resident and running, whatever order the hardware dispatches blocks in (with blockIdx as the tile id
  // -- -DGSA_PASS_DYNAMIC_TILES=0, 0.4-1.3 % faster -- forward progress would rest on in-order dispatch, which CUDA
  // does not promise, least of all with several builds sharing the device).
#if GSA_PASS_DYNAMIC_TILES
  const u32 tile = s_tile;
#else
  const u32 tile = blockIdx.x;
#endif
  const u32 tile_base = tile * (u32)TILE;
  const u32 valid = min((u32)TILE, a.n - tile_base);
  if (!GEN && a.pf_dist != 0u && tid == 0) {
    // The tile that the CTA taking this SM slot after me will most likely get: its 48 KB move from HBM to L2 while
    // the tiles in between are processed, so its loads meet L2 latency, not DRAM latency.
    const u64 pb = (u64)tile_base + (u64)a.pf_dist * TILE;
    if (pb < a.n) {
      const u32 cnt = (u32)min((u64)TILE, (u64)a.n - pb);
      l2_prefetch_bulk(static_cast<const KIN *>(a.keys_in) + pb, (cnt * (u32)sizeof(KIN)) & ~15u);
      l2_prefetch_bulk(a.vals_in + pb, (cnt * 4u) & ~15u);
    }
  }

  // ---- load (warp-striped: element order inside the tile is (warp, k, lane)) ----------
  u64 key[IPT];
  u32 val[IPT];
  const u32 wbase = tile_base + (u32)warp * (32u * IPT) + (u32)lane;
  // (A variant without the per-element bounds checks for full tiles -- 6 instructions per element less in the load and
  // in the scatter -- was SLOWER: 0.60 of the copy bandwidth with either, 0.57 with both, against 0.635: the compiler then
  // issues the 32 loads / 32 stores of a thread back to back, and the bursts delay the other CTAs' memory operations.)
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    const u32 idx = wbase + (u32)k * 32u;
    if (idx < a.n) {
      if (GEN) {
        gen_key0(a.gen, idx, key[k], val[k]);
      } else {
        if (sizeof(KIN) == 8) key[k] = ld_stream_u64(static_cast<const u64 *>(a.keys_in) + idx);
        else key[k] = ld_stream_u32(static_cast<const u32 *>(a.keys_in) + idx);
        val[k] = ld_stream_u32(a.vals_in + idx);
      }
    } else {
      key[k] = ~0ull;  // digit 255 at every shift; sits behind every real element of the tile
      val[k] = 0;
    }
  }

  // ---- per-warp digit counts and warp-local stable ranks in one step -----------------------
  // For each row k the lanes holding the same digit elect their lowest lane, which claims
  // `count` slots of the warp's bin (GSA_CLAIM: a plain u16 load + store, the leaders' digits are distinct) and hands the
  // base to its peers; local rank = base + number of peers in lower lanes.  Rows are issued in
  // order by the converged warp, so equal digits keep their (k, lane) order: stable.
  // Peer masks come from match.any when the first row shows few distinct digits in the warp
  // (match.any's cost grows with the number of distinct values) and from 8 ballots otherwise.
  u32 lrank[IPT / 2];  // two u16 per register
  const u32 lt = lanemask_lt();
  u16 *wh = whist + warp * RADIX;
  [[maybe_unused]] u32 *wh32 = reinterpret_cast<u32 *>(wh);  // (GSA_RANK_ATOMS form: the atomic works on the 32-bit word holding the bin's half)
  bool use_match;
  {
    const u32 d = (u32)(key[0] >> a.shift) & 255u;
    const u32 peers = peers_by_ballot(d);
    const u32 below = __popc(peers & lt);
    u32 base = 0;
    GSA_CLAIM(base, wh, wh32, d, peers, below);
    base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
    lrank[0] = base + below;
    use_match = __popc(__ballot_sync(0xffffffffu, below == 0)) <= 6;
  }
  if (use_match) {
#pragma unroll
    for (int k = 1; k < IPT; ++k) {
      const u32 d = (u32)(key[k] >> a.shift) & 255u;
      const u32 peers = __match_any_sync(0xffffffffu, d);
      const u32 below = __popc(peers & lt);
      u32 base = 0;
      GSA_CLAIM(base, wh, wh32, d, peers, below);
      base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
      if (k & 1) lrank[k >> 1] |= (base + below) << 16; else lrank[k >> 1] = base + below;
    }
  } else {
#pragma unroll
    for (int k = 1; k < IPT; ++k) {
      const u32 d = (u32)(key[k] >> a.shift) & 255u;
      const u32 peers = peers_by_ballot(d);
      const u32 below = __popc(peers & lt);
      u32 base = 0;
      GSA_CLAIM(base, wh, wh32, d, peers, below);
      base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
      if (k & 1) lrank[k >> 1] |= (base + below) << 16; else lrank[k >> 1] = base + below;
    }
  }
  __syncthreads();

  // ---- bin totals, per-warp offsets, exclusive scan over bins; publish the aggregate ----
  u32 cnt = 0, pub = 0, bin_ex = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const u32 c = whist[w * RADIX + tid];
      whist[w * RADIX + tid] = (u16)cnt;
      cnt += c;
    }
    pub = cnt - ((tid == RADIX - 1) ? ((u32)TILE - valid) : 0u);  // padding is not data
    if (tile == 0)
      st_volatile_u32(a.status + tid, st_pre(pub));
    else
      st_volatile_u32(a.status + (size_t)tile * RADIX + tid, st_agg(pub));
    u32 x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_wsum[warp] = x;
    cnt = x - cnt;  // exclusive within the warp
  }
  __syncthreads();
  if (tid < RADIX) {
    u32 add = 0;
    for (int i = 0; i < warp; ++i) add += s_wsum[i];
    const u32 ex = bin_ex = cnt + add;  // tile-local exclusive offset of the bin (= the offset of warp 0 after this step)
#pragma unroll
    for (int w = 0; w < WARPS; ++w) whist[w * RADIX + tid] = (u16)(whist[w * RADIX + tid] + ex);
  }
  __syncthreads();

  // ---- final rank inside the tile; stage in shared memory ---------------------------------
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    const u32 d = (u32)(key[k] >> a.shift) & 255u;
    const u32 r = wh[d] + ((k & 1) ? (lrank[k >> 1] >> 16) : (lrank[k >> 1] & 0xffffu));
    skeys[r] = (KOUT)key[k];
    svals[r] = val[k];
  }

  // ---- decoupled look-back: one thread per bin, four predecessors per round trip ------------
  if (tid < RADIX) {
    u32 excl = 0;
    if (tile != 0) {
      const u32 *base = a.status + tid;
      i64 t = (i64)tile - 1;
      const u32 done0 = st_pre(0);  // virtual tiles in front of tile 0
      // (Measured and dropped, profiles/r2/README.md: 8 / 16 tiles per far step, u16 aggregates with a 16 / 24 / 32 tile
      // window, prefixes from a dedicated scanner CTA, one-lane polling of not-ready rows -- every one slower than this.)
      for (;;) {
        u32 s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = (t - j >= 0) ? ld_volatile_u32(base + (size_t)(t - j) * RADIX) : done0;
        int used = 0;
        bool fin = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!fin && used == j && s[j] != 0u) {
            excl += (s[j] & 0x7fffffffu) - 1u;
            used = j + 1;
            fin = (s[j] & 0x80000000u) != 0u;
          }
        }
        if (fin) break;
        t -= used;  // used == 0: the nearest predecessor is not ready yet, poll again
      }
      st_volatile_u32(a.status + (size_t)tile * RADIX + tid, st_pre(excl + pub));
    }
    bin_gofs[tid] = a.bin_base[tid] + excl - bin_ex;
  }
  __syncthreads();

  // ---- scatter: consecutive threads write consecutive slots of a digit run -----------------
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    const u32 e = (u32)k * THREADS + (u32)tid;
    if (e < valid) {
      const KOUT kx = skeys[e];
      const u32 o = bin_gofs[(u32)((u64)kx >> a.shift) & 255u] + e;
      static_cast<KOUT *>(a.keys_out)[o] = kx;
      a.vals_out[o] = svals[e];
    }
  }
}

// ---------------------------------------------------------------------------------
// k_radix_pass_p: the same onesweep pass as a PERSISTENT kernel whose tiles arrive by bulk
// asynchronous copies (cp.async.bulk, the 1-D form of TMA; SASS UBLKCP) into a two-deep ring
// of shared-memory buffers:
//   * a CTA takes tiles by ticket (atomic counter), so the look-back only ever waits on tiles
//     that are held by CTAs that are resident and running -- forward progress does not depend
//     on the order in which the hardware dispatches blocks;
//   * while tile t is ranked, staged and scattered, the 48 KB of tile t' (the CTA's next
//     ticket) are already on their way from HBM, signalled through an mbarrier; no warp ever
//     sits on a global load of keys or values, and the registers hold one tile only while it
//     is being ranked;
//   * the buffer a tile was loaded into is reused as its staging area for the scatter.
// Ticket t + 2 is requested (result unused until the next trip) while t is processed, so the
// atomic's round trip is off the critical path as well.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
  u32 done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <int THREADS, int IPT>
struct PassPCfg {
  static constexpr int WARPS = THREADS / 32;
  static constexpr int TILE = THREADS * IPT;
  static constexpr size_t BUF = (size_t)TILE * 12;  // keys [TILE] then values [TILE]
  // two tile buffers, per-warp digit counts (u16), bin offsets, two mbarriers
  static constexpr size_t SMEM = 2 * BUF + (size_t)WARPS * RADIX * 2 + RADIX * 4 + 16;
};

template <int THREADS, int IPT, bool GEN, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_radix_pass_p(const PassArgs a, const u32 tiles) {
  static_assert(THREADS >= RADIX && THREADS % 32 == 0, "one thread per bin is assumed");
  static_assert(IPT % 2 == 0 && THREADS * IPT <= 65535, "ranks, counts and tile offsets are kept as u16");
  constexpr int WARPS = THREADS / 32;
  constexpr int TILE = THREADS * IPT;
  constexpr size_t BUF = PassPCfg<THREADS, IPT>::BUF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u16 *whist = reinterpret_cast<u16 *>(smem_raw + 2 * BUF);           // [WARPS][256] counts -> tile offsets
  u32 *bin_gofs = reinterpret_cast<u32 *>(whist + WARPS * RADIX);     // [256] global offset - local offset
  u64 *mbar = reinterpret_cast<u64 *>(bin_gofs + RADIX);              // [2]
  __shared__ u32 s_wsum[RADIX / 32];
  __shared__ u32 s_ticket[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = lanemask_lt();
  u16 *wh = whist + warp * RADIX;
  [[maybe_unused]] u32 *wh32 = reinterpret_cast<u32 *>(wh);

  // thread 0: the loader.  issue(t, b): bulk copies of tile t into buffer b, completion on mbar[b]
  auto issue = [&](u32 t, int b) {
    const u32 base = t * (u32)TILE;
    const u32 cnt = min((u32)TILE, a.n - base);
    const u32 kbytes = (cnt * 8u + 15u) & ~15u, vbytes = (cnt * 4u + 15u) & ~15u;  // the arrays are padded to 256 B
    unsigned char *buf = smem_raw + (size_t)b * BUF;
    mbar_expect_tx(&mbar[b], kbytes + vbytes);
    bulk_load(buf, static_cast<const u64 *>(a.keys_in) + base, kbytes, &mbar[b]);
    bulk_load(buf + (size_t)TILE * 8, a.vals_in + base, vbytes, &mbar[b]);
  };

  u32 ahead = 0;  // thread 0: the ticket after the next one (requested early, consumed one trip later)
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
    s_ticket[0] = atomicAdd(a.counter, 1u);
  }
  __syncthreads();
  u32 tile = s_ticket[0];
  int cur = 0;
  u32 parity = 0;  // bit b = phase of mbar[b] to wait for
  if (tid == 0) {
    if (!GEN && tile < tiles) issue(tile, 0);
    ahead = atomicAdd(a.counter, 1u);
  }
  while (tile < tiles) {
    if (tid == 0) {
      const u32 nt = ahead;
      s_ticket[cur ^ 1] = nt;
      if (!GEN && nt < tiles) issue(nt, cur ^ 1);
      ahead = (nt < tiles) ? atomicAdd(a.counter, 1u) : nt;
    }
    for (int i = tid; i < WARPS * RADIX / 2; i += THREADS) reinterpret_cast<u32 *>(whist)[i] = 0;
    const u32 tile_base = tile * (u32)TILE;
    const u32 valid = min((u32)TILE, a.n - tile_base);
    u64 *skeys = reinterpret_cast<u64 *>(smem_raw + (size_t)cur * BUF);
    u32 *svals = reinterpret_cast<u32 *>(skeys + TILE);

    // ---- the tile: element order is (warp, k, lane) ------------------------------------------
    u64 key[IPT];
    u32 val[IPT];
    if (!GEN) {
      mbar_wait(&mbar[cur], (parity >> cur) & 1u);
      parity ^= 1u << cur;
    }
    const u32 lbase = (u32)warp * (32u * IPT) + (u32)lane;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const u32 e = lbase + (u32)k * 32u;
      if (e < valid) {
        if (GEN) {
          gen_key0(a.gen, tile_base + e, key[k], val[k]);
        } else {
          key[k] = skeys[e];
          val[k] = svals[e];
        }
      } else {
        key[k] = ~0ull;  // digit 255 at every shift; sits behind every real element of the tile
        val[k] = 0;
      }
    }
    __syncthreads();  // counts are zero; everybody holds its elements: the buffer becomes the staging area

    // ---- per-warp digit counts and warp-local stable ranks (see k_radix_pass) -----------------
    u32 lrank[IPT / 2];
    bool use_match;
    {
      const u32 d = (u32)(key[0] >> a.shift) & 255u;
      const u32 peers = peers_by_ballot(d);
      const u32 below = __popc(peers & lt);
      u32 base = 0;
      GSA_CLAIM(base, wh, wh32, d, peers, below);
      base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
      lrank[0] = base + below;
      use_match = __popc(__ballot_sync(0xffffffffu, below == 0)) <= 6;
    }
    if (use_match) {
#pragma unroll
      for (int k = 1; k < IPT; ++k) {
        const u32 d = (u32)(key[k] >> a.shift) & 255u;
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const u32 below = __popc(peers & lt);
        u32 base = 0;
        GSA_CLAIM(base, wh, wh32, d, peers, below);
        base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
        if (k & 1) lrank[k >> 1] |= (base + below) << 16; else lrank[k >> 1] = base + below;
      }
    } else {
#pragma unroll
      for (int k = 1; k < IPT; ++k) {
        const u32 d = (u32)(key[k] >> a.shift) & 255u;
        const u32 peers = peers_by_ballot(d);
        const u32 below = __popc(peers & lt);
        u32 base = 0;
        GSA_CLAIM(base, wh, wh32, d, peers, below);
        base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
        if (k & 1) lrank[k >> 1] |= (base + below) << 16; else lrank[k >> 1] = base + below;
      }
    }
    __syncthreads();

    // ---- bin totals, per-warp offsets, exclusive scan over bins; publish the aggregate ----
    u32 cnt = 0, pub = 0, bin_ex = 0;
    if (tid < RADIX) {
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const u32 c = whist[w * RADIX + tid];
        whist[w * RADIX + tid] = (u16)cnt;
        cnt += c;
      }
      pub = cnt - ((tid == RADIX - 1) ? ((u32)TILE - valid) : 0u);  // padding is not data
      if (tile == 0)
        st_volatile_u32(a.status + tid, st_pre(pub));
      else
        st_volatile_u32(a.status + (size_t)tile * RADIX + tid, st_agg(pub));
      u32 x = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) s_wsum[warp] = x;
      cnt = x - cnt;  // exclusive within the warp
    }
    __syncthreads();
    if (tid < RADIX) {
      u32 add = 0;
      for (int i = 0; i < warp; ++i) add += s_wsum[i];
      const u32 ex = bin_ex = cnt + add;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) whist[w * RADIX + tid] = (u16)(whist[w * RADIX + tid] + ex);
    }
    __syncthreads();

    // ---- final rank inside the tile; stage in the tile's own buffer -----------------------------
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const u32 d = (u32)(key[k] >> a.shift) & 255u;
      const u32 r = wh[d] + ((k & 1) ? (lrank[k >> 1] >> 16) : (lrank[k >> 1] & 0xffffu));
      skeys[r] = key[k];
      svals[r] = val[k];
    }

    // ---- decoupled look-back: one thread per bin, four predecessors per round trip ------------
    if (tid < RADIX) {
      u32 excl = 0;
      if (tile != 0) {
        const u32 *base = a.status + tid;
        i64 t = (i64)tile - 1;
        const u32 done0 = st_pre(0);
        for (;;) {
          u32 sv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) sv[j] = (t - j >= 0) ? ld_volatile_u32(base + (size_t)(t - j) * RADIX) : done0;
          int used = 0;
          bool fin = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!fin && used == j && sv[j] != 0u) {
              excl += (sv[j] & 0x7fffffffu) - 1u;
              used = j + 1;
              fin = (sv[j] & 0x80000000u) != 0u;
            }
          }
          if (fin) break;
          t -= used;
        }
        st_volatile_u32(a.status + (size_t)tile * RADIX + tid, st_pre(excl + pub));
      }
      bin_gofs[tid] = a.bin_base[tid] + excl - bin_ex;
    }
    __syncthreads();

    // ---- scatter: consecutive threads write consecutive slots of a digit run -----------------
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const u32 e = (u32)k * THREADS + (u32)tid;
      if (e < valid) {
        const u64 kx = skeys[e];
        const u32 o = bin_gofs[(u32)(kx >> a.shift) & 255u] + e;
        static_cast<u64 *>(a.keys_out)[o] = kx;
        a.vals_out[o] = svals[e];
      }
    }
    fence_proxy_async();  // this buffer is the target of a bulk copy two trips from now
    __syncthreads();
    tile = s_ticket[cur ^ 1];
    cur ^= 1;
  }
}
#endif  // __CUDACC__

}  // namespace gsa
