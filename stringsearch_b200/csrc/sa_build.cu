// sa_build.cu -- suffix-array construction on the GPU: prefix doubling with
// singleton-group discard.  Replaces the *result* of divsufsort()
// (reference: crates/cdivsufsort/c-sources/divsufsort.c:331-370,
// crates/divsufsort/src/divsufsort.rs:3-37); none of the induced-sorting code is
// reproduced.
//
// Round 0   text bytes -> dense codes (b bits) -> bit-packed stream; key(i) = the next
//           k = floor(64/b) symbols of suffix i, zero padded; LSD radix sort of (key, i);
//           adjacent-key-differs flags + head/tail scans give every group's slot range.
// Round r   live (not yet unique) suffixes only, walked in text order so that the two rank
//           reads per suffix are coalesced: key = (rank[i], rank[i + h]); sort, re-flag,
//           finalise singletons into SA, compact the surviving slots.
//
// rank[i] is a *label* of i's group: some SA slot inside the group's slot range, plus one
// (0 = "past the end of the text", bit 31 = suffix finalised).  Ranges of different groups
// are disjoint and ordered, so labels order groups correctly, and a group KEEPS its label
// for as long as the label stays inside its shrinking range.  Only suffixes whose group
// gets a new label (the middle of the new range) or that become unique are written -- the
// random 4-byte scatter is the most expensive memory operation of a round (22-27 G/s on
// B200 whatever the store flavour, tools/ubench/gather_scatter.cu), and with slot-of-head
// ranks nearly every suffix of a repetitive text would be rewritten in every round.
//
// Data layout in HBM (n = text length, L = live suffixes, all arrays contiguous):
//   packed  u64[n*b/64 + 2]   keys u64[n] x2   vals(suffix) u32[n] x2
//   pos u32[n] x2 (SA slot of each live element, SA order)
//   lst u32[n] x2 (live suffixes, text order)   rank u32[n]   SA i32[n] (caller's)
//   bag_sufx / bag_pos u32[n] x2 (members of tiny groups)   G u64[n] + tables per n/256 labels (huge groups)
//   + histograms, pass look-back status, per-tile head / tail carries of the rebuild and slot kernels.
#include "builder.h"
#include <stdlib.h>

#include <algorithm>
#include <atomic>
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
  if ((64u % b) == 0u && i + 64u / b <= n && (reinterpret_cast<uintptr_t>(T) & 7u) == 0u) {
    // whole symbols per word (b = 1, 2, 4, 8) and all of them inside the text: the word's 64 / b source bytes start at a
    // multiple of 8 -- eight bytes per load instead of one (byte loads ran this kernel at 1.3 TB/s)
    const u32 spw = 64u / b;
    const u64 *src = reinterpret_cast<const u64 *>(T + i);  // T is 8-byte aligned (checked), i a multiple of 8
    for (u32 s8 = 0; s8 < spw; s8 += 8u) {
      const u64 x = ld_stream_u64(src + (s8 >> 3));
#pragma unroll
      for (int k = 0; k < 8; ++k) word = (word << b) | (u64)scode[(u32)(x >> (8 * k)) & 255u];
    }
    packed[q] = word;
    return;
  }
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

// When a radix digit covers whole symbols (8 % b == 0, key_bits % 8 == 0) every digit of every
// round-0 key is one of the n 8-bit windows W[j] = stream bits [j*b, j*b + 8): digit p of
// suffix i is W[i + s_p] with s_p = (key_bits - 8(p+1)) / b.  So one 256-bin histogram of the
// windows replaces the npass histograms (one shared-memory atomic per suffix instead of
// npass), up to the few windows at the two ends that k_hist0_finish corrects.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_hist0_windows(const u64 *__restrict__ packed, u32 n, u32 b, u32 *__restrict__ hfull) {
  __shared__ u32 shist[RADIX];
  for (int i = threadIdx.x; i < RADIX; i += THREADS) shist[i] = 0;
  __syncthreads();
  // one packed word (64 / b windows) per thread and step: two loads per word instead of two per window; runs of equal
  // windows (a^n ...) are added once
  const u32 spw = 64u / b;
  const u32 nw = (n + spw - 1u) / spw;
  for (u32 q = blockIdx.x * THREADS + threadIdx.x; q < nw; q += gridDim.x * THREADS) {
    const u64 w0 = __ldg(packed + q), w1 = __ldg(packed + q + 1);  // (the stream has two spare zero words)
    const u32 cnt = min(spw, n - q * spw);
    u32 prev = 0, run = 0;
    for (u32 s = 0; s < cnt; ++s) {
      const u32 sh = s * b;
      const u32 win = (u32)((sh ? ((w0 << sh) | (w1 >> (64u - sh))) : w0) >> 56);
      if (run != 0u && win == prev) {
        ++run;
      } else {
        if (run != 0u) atomicAdd(&shist[prev], run);
        prev = win;
        run = 1u;
      }
    }
    if (run != 0u) atomicAdd(&shist[prev], run);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RADIX; i += THREADS)
    if (shist[i]) atomicAdd(&hfull[i], shist[i]);
}

__global__ void __launch_bounds__(RADIX) k_hist0_finish(const u64 *__restrict__ packed, u32 n, u32 b, u32 key_bits, int npass,
                                                        const u32 *__restrict__ hfull, u32 *__restrict__ ghist) {
  const u32 v = threadIdx.x;
  const u32 full = hfull[v];
  for (int p = 0; p < npass; ++p) ghist[p * RADIX + v] = full;
  __syncthreads();
  if (v == 0) {
    for (int p = 0; p < npass; ++p) {
      const u32 sp = (key_bits - 8u * (u32)(p + 1)) / b;  // suffix i reads window i + sp
      // windows 0 .. sp-1 belong to no suffix; windows n .. n+sp-1 (all zero) do
      for (u32 j = 0; j < sp && j < n; ++j) {
        const u64 bit = (u64)j * b;
        const u32 sh = (u32)bit & 63u;
        const u64 w0 = packed[bit >> 6], w1 = packed[(bit >> 6) + 1];
        const u32 win = (u32)((sh ? ((w0 << sh) | (w1 >> (64u - sh))) : w0) >> 56);
        ghist[p * RADIX + win] -= 1u;
      }
      ghist[p * RADIX + 0] += min(sp, n);
    }
  }
}

// ------------------------------------------------------------------------------------
// Repetitiveness probe: how many of `m` pseudo-random suffixes share their first
// key_bits / b symbols with another sampled suffix?  (Open-addressing insert into a small
// hash table; a hit on an equal key counts as a duplicate.)  Text without long repeats gives
// ~0; repetitive text gives ~m.  The host uses it to pick the round-0 key depth.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sample_dups(const KeyGen g, u32 m, u64 *__restrict__ table, u32 table_mask,
                                                     u32 *__restrict__ dups) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  u32 x = s * 2654435761u + 12345u;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  const u32 i = (u32)(((u64)x * (u64)g.n) >> 32);
  const u64 key = key_of_suffix(g, i) + 1ull;  // 0 marks an empty table slot
  u32 slot = (u32)((key * 0x9E3779B97F4A7C15ull) >> 40) & table_mask;
  for (int probe = 0; probe < 16; ++probe) {
    const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(table + slot), 0ull, (unsigned long long)key);
    if (old == 0ull) return;
    if (old == key) { atomicAdd(dups, 1u); return; }
    slot = (slot + 1u) & table_mask;
  }
}

// ------------------------------------------------------------------------------------
// Rounds >= 1, step 1: walk the candidate suffixes in text order, drop the finalised ones,
// build the sort key of the live ones, histogram its digits, and emit
//   (key, suffix) -> sort input,   suffix -> next round's candidate list.
//   key = label(i) << lab_bits | label(i + h)     (0 when i + h is past the end: a proper
//                                                  prefix sorts first)
// Consecutive candidates are consecutive in the text (up to a permutation of 4096-element
// chunks), so rank[i] and rank[i + h] are streamed, not gathered.  Output slots come from one
// atomicAdd per chunk; the order of the sort input does not matter.
// ------------------------------------------------------------------------------------
constexpr u32 NO_TAIL = 0xffffffffu;
constexpr u32 NO_TAIL_IDX = 0xffffffffu;
constexpr u32 RANK_DEAD = 0x80000000u;
constexpr u32 RANK_MASK = 0x7fffffffu;

// ---- huge groups ---------------------------------------------------------------------------
// In repetitive text a few groups hold almost every live suffix, and in each round almost all
// members of such a group share one sort key (they pair with the same group h symbols on).
// Sorting them is pointless: they stay one group.  So for a HUGE group (>= HUGE_T slots) one
// second key half rho* is chosen per round -- the one of a representative member, which is
// refreshed from the members that were inert in the previous round (they are exactly the
// members of the group now) -- and members whose label(i + h) equals it are INERT: they stay
// in the live list, keep their label, and are neither sorted nor rewritten.  Any choice of
// rho* is correct; a bad one only means a round in which the group is sorted in full.  Only the other members are sorted; those below rho* fill the group's slot range
// from the left, those above from the right, and the range of the inert block shrinks
// accordingly in the per-group table G[label] = (first slot, last slot).
// A label that is a multiple of HUGE_M marks a huge group; small groups never use such labels,
// so a reader knows from the label alone whether the group's state lives in the tables.  What
// happens to a huge group between rounds (label fell out of the shrunken range, group became
// small, or unique) is decided once per round by k_huge_prepare and published in STATE; every
// reader of a huge label resolves it through STATE, members rewrite their own rank on the fly.
constexpr u32 HUGE_M = 256;
// Smallest group that is handled through the group tables.  The tables are indexed by label / HUGE_M, so
// 2 * HUGE_M is the floor.  Measured (tools/shapes_bench.py, 256 MiB): 65536 -> 512 changes nothing on
// rep_1G, zeros, period-7 text, but a text whose groups have a few thousand members (period 100 000,
// 2684 repeats) drops from 308 ms to 86 ms: those groups were sorted in full in every round.
#ifndef GSA_HUGE_T
#define GSA_HUGE_T 512
#endif
constexpr u32 HUGE_T = GSA_HUGE_T;
static_assert(HUGE_T >= 2 * HUGE_M, "a huge range must hold two candidate labels");
constexpr u32 STATE_FINAL = 0x80000000u;
constexpr u32 HUGE_REPS = 16;  // representatives per huge group: rho* is the plurality of their second key halves

__device__ __forceinline__ bool is_huge_label(u32 lab) { return (lab & (HUGE_M - 1u)) == 0u; }
// Representatives are kept as 64-bit keys  (255 - round) << 56 | hash << 32 | suffix  and updated with
// atomicMin: a slot holds the member with the smallest hash among the volunteers of the newest round.
// That makes the choice independent of the order in which blocks run (with plain stores the last
// writer wins, i.e. always a member from the end of the text -- on a^n exactly the members that
// turn into the minority one round later).  All ones = empty.
__device__ __forceinline__ u64 rep_key(u32 round, u32 hash24, u32 sufx) {
  return ((u64)(255u - round) << 56) | ((u64)(hash24 & 0xffffffu) << 32) | sufx;
}
__device__ __forceinline__ u32 mix32(u32 x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// The BAG.  A suffix whose group has at most TINY_MAX members leaves the text-order walk and
// the radix sort for good: such groups are kept, group by group, as (suffix, slot) entries in
// a list and refined in place every round -- gather label(i + h), sort inside the group
// (k_bag_refine, one thread per entry), split, finalise.  For these suffixes a round costs one
// random 4-byte gather instead of 8 radix passes over 12 bytes.
// The class of a group is readable from its label alone:
//   multiple of HUGE_M : huge   (>= HUGE_T slots), state in the group tables
//   other even label   : medium, sorted through the text-order path every round
//   odd label          : tiny, lives in the bag; always the canonical label of its slot range
// (tiny_max == 0 switches the tiny class off: sparse mode, which keeps no label for most suffixes.)
constexpr u32 TINY_MAX = 32;
constexpr u32 BAG_HEAD = 0x80000000u;  // bag entry: first member of its group

__device__ __forceinline__ u32 tiny_label(u32 s) { return (s & 1u) ? s + 2u : s + 1u; }  // first odd label in [s+1, ..]

// Label (1-based slot) for the group occupying slots [s, e], e > s.
// `avoid`: a label that must not be chosen (the label of the huge group this one is split
// from: its table entry still belongs to the inert block); 0 = none.
__device__ __forceinline__ u32 pick_label(u32 s, u32 e, u32 avoid, u32 tiny_max) {
  const u32 size = e - s + 1u;
  const u32 mid = s + ((e - s) >> 1) + 1u;
  if (size >= HUGE_T) {
    u32 lab = mid & ~(HUGE_M - 1u);
    if (lab < s + 1u) lab += HUGE_M;
    if (lab == avoid) lab = (lab + HUGE_M <= e + 1u) ? lab + HUGE_M : lab - HUGE_M;
    return lab;
  }
  if (size <= tiny_max) return tiny_label(s);
  u32 lab = mid;
  if (tiny_max == 0u) {  // any label that is no multiple of HUGE_M
    if (is_huge_label(lab)) lab = (lab + 1u <= e + 1u) ? lab + 1u : lab - 1u;
    return lab;
  }
  if (lab & 1u) lab = (lab + 1u <= e + 1u) ? lab + 1u : lab - 1u;      // size > TINY_MAX: room on both sides
  if (is_huge_label(lab)) lab = (lab + 2u <= e + 1u) ? lab + 2u : lab - 2u;
  return lab;
}

// rank word -> current label; *fin = the suffix is (or has just become) unique.  `st` is the
// STATE entry of the word's label (only looked at for live huge labels), so that callers can
// issue the table loads of many words before resolving any of them.
__device__ __forceinline__ bool needs_state(u32 w) { return !(w & RANK_DEAD) && w != 0u && is_huge_label(w); }
__device__ __forceinline__ u32 resolve_with(u32 w, u64 st, u32 round, bool *fin) {
  *fin = false;
  if (w & RANK_DEAD) { *fin = true; return w & RANK_MASK; }
  if (!is_huge_label(w)) return w;
  if ((u32)(st >> 32) != round) return w;  // nothing happened to this group
  const u32 code = (u32)st;
  if (code & STATE_FINAL) { *fin = true; return code & RANK_MASK; }
  return code;  // moved to a new label
}
__device__ __forceinline__ u32 resolve_label(u32 w, const u64 *__restrict__ state, u32 round, bool *fin) {
  const u64 st = needs_state(w) ? __ldg(state + (w / HUGE_M)) : 0ull;
  return resolve_with(w, st, round, fin);
}

// Round 0 creates the huge groups of a repetitive text: scattering a label to each of their
// members would be n random 4-byte writes.  Instead the group head registers
// (round-0 key -> label) in a small hash table and k_rank_huge0 walks the text once, in order:
// a suffix whose key is in the table gets its label by a streaming write.
struct HugeKeyTable {
  u64 *keys;    // [cap]
  u32 *labels;  // [cap] 0 = empty
  u32 mask;     // cap - 1
};
__device__ __forceinline__ u32 hkt_hash(u64 key, u32 mask) { return (u32)((key * 0x9E3779B97F4A7C15ull) >> 40) & mask; }
__device__ __forceinline__ void hkt_insert(const HugeKeyTable &t, u64 key, u32 label) {
  u32 sl = hkt_hash(key, t.mask);
  for (;;) {
    if (atomicCAS(t.labels + sl, 0u, label) == 0u) { t.keys[sl] = key; return; }
    sl = (sl + 1u) & t.mask;
  }
}
__global__ void __launch_bounds__(256) k_rank_huge0(const KeyGen g, const HugeKeyTable t, u32 first_short, u32 *__restrict__ rank) {
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < first_short; i += stride) {
    const u64 key = key_of_suffix(g, i);
    u32 sl = hkt_hash(key, t.mask);
    for (;;) {
      const u32 lab = __ldg(t.labels + sl);
      if (lab == 0u) break;
      if (__ldg(t.keys + sl) == key) { rank[i] = lab; break; }
      sl = (sl + 1u) & t.mask;
    }
  }
}

// Once per round, one thread per huge group: classify it, publish the verdict in STATE for the
// readers of this round, and rebuild the list of huge groups.
__global__ void __launch_bounds__(256) k_huge_prepare(const u32 *__restrict__ hin, u32 cnt, u64 *__restrict__ G,
                                                      u64 *__restrict__ state, u64 *__restrict__ rep, u32 round,
                                                      u32 tiny_max, u32 *__restrict__ hout, u32 *__restrict__ hout_count,
                                                      u32 *__restrict__ verdicts, u32 *__restrict__ seen) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cnt) return;
  const u32 lab = hin[j];
  // a group that was sorted in full (no rho*) and did not split is appended again by the rebuild: keep one entry
  if (seen != nullptr && atomicExch(seen + lab / HUGE_M, round) == round) return;
  const u64 g = G[lab];
  const i64 gs = (i64)(i32)(u32)g, ge = (i64)(i32)(u32)(g >> 32);
  const i64 size = ge - gs + 1;
  const u64 tag = (u64)round << 32;
  if (size <= 0) return;  // every member left the inert block
  if (size == 1) { state[lab / HUGE_M] = tag | STATE_FINAL | (u32)(gs + 1); *verdicts = 1u; return; }
  if (size < (i64)HUGE_T || !((i64)lab >= gs + 1 && (i64)lab <= ge + 1)) {
    // new huge label inside the range, or the label of the class the group has shrunk to (if tiny,
    // its members -- still in the text-order list -- are all sorted this round and the rebuild
    // moves them to the bag)
    const u32 nl = pick_label((u32)gs, (u32)ge, 0u, tiny_max);
    G[nl] = g;
    state[lab / HUGE_M] = tag | nl;
    *verdicts = 1u;  // readers must consult STATE this round
    if (is_huge_label(nl)) {
      for (u32 x = 0; x < HUGE_REPS; ++x) rep[(nl / HUGE_M) * HUGE_REPS + x] = rep[(lab / HUGE_M) * HUGE_REPS + x];
      hout[atomicAdd(hout_count, 1u)] = nl;
    }
    return;
  }
  hout[atomicAdd(hout_count, 1u)] = lab;
}

// After the verdicts: rho* of every huge group for this round = the most frequent label(rep + h)
// among its representatives that still belong to the group (none left: the group goes
// unfiltered this round).  One representative would do for correctness, but its key half is the
// dominant one only with the probability that a member is inert, and a miss costs a full sort
// of the group.
__global__ void __launch_bounds__(256) k_huge_rho(const u32 *__restrict__ hl, const u32 *__restrict__ hl_count,
                                                  const u32 *__restrict__ rank, const u64 *__restrict__ state,
                                                  const u64 *__restrict__ rep, u64 *__restrict__ rho, u32 round, u64 h,
                                                  u32 n) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *hl_count) return;
  const u32 lab = hl[j];
  u32 val[HUGE_REPS];
  u32 nv = 0;
  for (u32 x = 0; x < HUGE_REPS; ++x) {
    const u32 r = (u32)rep[(lab / HUGE_M) * HUGE_REPS + x];
    if (r >= n) continue;
    bool fin;
    const u32 w = rank[r];
    if (w == 0u || resolve_label(w, state, round, &fin) != lab || fin) continue;
    const u64 t = (u64)r + h;
    u32 r2 = 0;
    if (t < n) {
      const u32 w2 = rank[t];
      if (w2 == 0u) continue;  // sparse mode keeps no label for it
      r2 = resolve_label(w2, state, round, &fin);
    }
    val[nv++] = r2;
  }
  if (nv == 0) return;
  u32 best = val[0], bestc = 0;
  for (u32 x = 0; x < nv; ++x) {
    u32 c = 0;
    for (u32 y = 0; y < nv; ++y) c += (val[y] == val[x]) ? 1u : 0u;
    if (c > bestc) { bestc = c; best = val[x]; }
  }
  rho[lab / HUGE_M] = ((u64)round << 32) | best;
}

struct GatherArgs {
  const u32 *lst_in;  // null: candidates are 0..Lin-1
  u32 Lin;
  const u32 *Lin_ptr;  // non-null: the number of candidates is read from device memory (k_prefilter's output)
  u32 *rank;          // read; members of re-labelled / finalised huge groups rewrite their own entry
  i32 *SA;
  u32 n;
  u64 h;
  u32 lab_bits;
  int npass;
  u32 round;
  int filter;         // leave inert members of huge groups out of the sort
  u32 tiny_max;       // > 0: suffixes with an odd label live in the bag and are dropped from the list
  const u64 *state;   // [n / HUGE_M + 2] verdicts of k_huge_prepare, tagged with the round
  const u32 *verdicts;  // [1] != 0: some huge group got a verdict this round (else STATE need not be read)
  const u64 *rho;     // [n / HUGE_M + 2] round << 32 | rho*  (k_huge_rho)
  u64 *rep;           // [n / HUGE_M + 2][HUGE_REPS] members of every huge group (rep_key), refreshed from the inert ones
  u64 *keys_out;
  u32 *vals_out;
  u32 *lst_out;
  u32 *counter;       // [0] live suffixes (lst_out), [1] suffixes to sort (keys_out / vals_out); zeroed before launch
  u32 *ghist;
  // sparse mode (few suffixes survived round 0): rank[] holds 0 for every suffix that was
  // already unique after round 0; its label is recomputed on demand from the round-0 order.
  const i32 *sa0;  // null unless sparse: SA after round 0, every slot filled
  KeyGen gen;
  u32 *todo;        // sparse: (output slot, suffix) pairs whose second key half is still missing
  u32 *todo_count;
};

// Label of a suffix that was unique after round 0 (sparse mode): its SA slot + 1, found by
// binary search of its round-0 key in the round-0 order.  Equal keys are ordered short
// suffixes first (shortest first), then the long ones, so a suffix sits behind every short
// suffix with the same zero-padded key (and, if short itself, only behind the shorter ones).
__device__ __forceinline__ u32 lazy_label(const KeyGen &g, const i32 *__restrict__ sa0, u32 t) {
  const u64 kt = key_of_suffix(g, t);
  u32 lo = 0, hi = g.n;
  while (lo < hi) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if (key_of_suffix(g, (u32)__ldg(sa0 + mid)) < kt) lo = mid + 1; else hi = mid;
  }
  u32 extra = 0;
  if ((kt & ((1ull << g.b) - 1ull)) == 0ull) {  // only a key ending in symbol 0 can equal a padded one
    const u32 first_short = g.n - g.ns;
    const u32 from = (t >= first_short) ? t + 1u : first_short;
    for (u32 j = from; j < g.n; ++j) extra += (key_of_suffix(g, j) == kt) ? 1u : 0u;
  }
  return lo + extra + 1u;
}

template <int THREADS, int IPT, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_gather(const GatherArgs a) {
  constexpr u32 CH = THREADS * IPT;
  __shared__ u32 shist[MAX_PASSES * RADIX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < a.npass * RADIX; i += THREADS) shist[i] = 0;
  __syncthreads();
  const u32 lt = lanemask_lt();
  const bool anyv = __ldg(a.verdicts) != 0u;
  const u32 Lin = a.Lin_ptr ? __ldg(a.Lin_ptr) : a.Lin;
  const u32 nchunks = (Lin + CH - 1) / CH;
  for (u32 chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    // warp w owns the contiguous sub-chunk [w*32*IPT, (w+1)*32*IPT); row k = 32 consecutive candidates
    const u32 wb = chunk * CH + (u32)warp * (32u * IPT) + (u32)lane;
    u32 sfx[IPT], w[IPT], r2[IPT];
    // three rounds of independent loads: candidates, their ranks, the ranks h further on
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const u32 c = wb + (u32)k * 32u;
      sfx[k] = (c < Lin) ? (a.lst_in ? __ldg(a.lst_in + c) : c) : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      w[k] = (sfx[k] != 0xffffffffu) ? __ldcg(a.rank + sfx[k]) : RANK_DEAD;
      if (w[k] == 0u) w[k] = RANK_DEAD;  // sparse mode: unique since round 0
      if (a.tiny_max && (w[k] & 1u)) w[k] = RANK_DEAD;  // tiny group: handled in the bag from now on
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const u64 t = (u64)sfx[k] + a.h;
      r2[k] = (!(w[k] & RANK_DEAD) && t < a.n) ? __ldcg(a.rank + t) : 0u;
    }
    // resolve huge labels through this round's verdicts; members fix their own entry.  The table
    // loads of all IPT rows are issued together, three times: STATE of the suffix, STATE of
    // suffix + h, rho* of the (resolved) group.
    u64 tb[IPT];
    if (anyv) {  // (warp-uniform) in most rounds no huge group changes its state
#pragma unroll
    for (int k = 0; k < IPT; ++k) tb[k] = needs_state(w[k]) ? __ldg(a.state + (w[k] / HUGE_M)) : 0ull;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      if (!(w[k] & RANK_DEAD)) {
        bool fin;
        const u32 lab = resolve_with(w[k], tb[k], a.round, &fin);
        if (fin) {  // its huge group has shrunk to this one suffix
          a.SA[lab - 1u] = (i32)sfx[k];
          a.rank[sfx[k]] = RANK_DEAD | lab;
          w[k] = RANK_DEAD;
        } else {
          if (lab != w[k]) a.rank[sfx[k]] = lab;
          w[k] = lab;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k) tb[k] = (!(w[k] & RANK_DEAD) && needs_state(r2[k])) ? __ldg(a.state + (r2[k] / HUGE_M)) : 0ull;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      bool fin;
      if (!(w[k] & RANK_DEAD) && r2[k] != 0u) r2[k] = resolve_with(r2[k], tb[k], a.round, &fin);
    }
    } else {
#pragma unroll
      for (int k = 0; k < IPT; ++k) r2[k] &= RANK_MASK;  // a finalised suffix: its slot + 1
    }
#pragma unroll
    for (int k = 0; k < IPT; ++k)
      tb[k] = (a.filter && !(w[k] & RANK_DEAD) && is_huge_label(w[k])) ? __ldg(a.rho + (w[k] / HUGE_M)) : 0ull;
    u32 offl[IPT], offs[IPT];  // slot offsets inside the warp's output runs (live list, sort input)
    u32 srt = 0;               // bit k: element k goes to the sort
    u32 wl = 0, ws = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const bool lv = (w[k] & RANK_DEAD) == 0u;
      bool so = lv;
      if (lv && a.filter && is_huge_label(w[k])) {
        const u64 e = tb[k];
        if ((u32)(e >> 32) == a.round && (u32)e == r2[k]) {
          so = false;  // inert: shares the group's dominant key
          // one in 256 of them volunteers as a representative of its group for the next round
          if (((sfx[k] * 0x9e3779b1u) >> 24) == (a.round & 0xffu)) {  // cheap 1-in-256 pre-filter
            const u32 hsh = mix32(sfx[k] ^ (a.round * 0x9e3779b9u));
            atomicMin(reinterpret_cast<unsigned long long *>(a.rep + (w[k] / HUGE_M) * HUGE_REPS + ((hsh >> 8) & (HUGE_REPS - 1u))),
                      (unsigned long long)rep_key(a.round, hsh >> 11, sfx[k]));
          }
        }
      }
      const u64 kx = so ? (((u64)w[k] << a.lab_bits) | r2[k]) : 0ull;
      const u32 ml = __ballot_sync(0xffffffffu, lv), ms = __ballot_sync(0xffffffffu, so);
      offl[k] = wl + (u32)__popc(ml & lt);
      offs[k] = ws + (u32)__popc(ms & lt);
      wl += (u32)__popc(ml);
      ws += (u32)__popc(ms);
      srt |= (so ? 1u : 0u) << k;
      if (ms) hist_add(shist, kx, so, a.npass);  // warp-uniform: rows of inert members skip it
    }
    // every warp reserves its own output ranges: the lists stay in text order within a warp's
    // 32 * IPT candidates (which is what keeps the next round's reads coalesced), and no warp
    // ever waits for another one's loads at a barrier
    // (both counters live in one 64-bit word -- live count low, sort count high -- so a warp needs one
    // atomic, and none that returns a value when it has nothing to place: every warp of the grid
    // hits this one address, 4 M times per round at 1 GiB)
    u32 bl = 0, bs = 0;
    if (lane == 0 && (wl | ws)) {
      unsigned long long *both = reinterpret_cast<unsigned long long *>(a.counter);
      const unsigned long long add = ((unsigned long long)ws << 32) | wl;
      if (ws != 0u || a.lst_out != nullptr) {
        const unsigned long long old = atomicAdd(both, add);
        bl = (u32)old;
        bs = (u32)(old >> 32);
      } else {
        atomicAdd(both, add);  // result unused: a reduction, nobody waits for it
      }
    }
    bl = __shfl_sync(0xffffffffu, bl, 0);
    bs = __shfl_sync(0xffffffffu, bs, 0);
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      if (a.lst_out != nullptr && (w[k] & RANK_DEAD) == 0u) a.lst_out[bl + offl[k]] = sfx[k];
      if ((srt >> k) & 1u) {
        const u32 o = bs + offs[k];
        a.keys_out[o] = ((u64)w[k] << a.lab_bits) | r2[k];
        a.vals_out[o] = sfx[k];
        if (a.sa0 != nullptr && r2[k] == 0u && (u64)sfx[k] + a.h < a.n) {
          // sparse mode: suffix i + h has been unique since round 0 and has no stored label;
          // k_lazy_fill computes it (one binary search per entry, all in flight at once)
          const u32 q = atomicAdd(a.todo_count, 1u);
          a.todo[2 * q] = o;
          a.todo[2 * q + 1] = (u32)(sfx[k] + a.h);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < a.npass * RADIX; i += THREADS)
    if (shist[i]) atomicAdd(&a.ghist[i], shist[i]);
}

// ------------------------------------------------------------------------------------
// Dense rounds (every text position is a candidate, most of them inert members of huge groups):
// a streaming pre-filter in front of k_gather.  Four consecutive suffixes per thread with 128-bit
// loads of rank[i..i+3] and rank[i+h..i+h+3]; a suffix that is
//   * finalised or in the bag                       -> dropped, as k_gather would,
//   * a member of a huge group that got no verdict this round, whose partner label got none
//     either and equals the group's rho*            -> INERT: counted as live, 1 in 256 volunteers
//                                                      as a representative, nothing else happens,
//   * anything else                                 -> appended to the candidate list that k_gather
//                                                      then walks with its full logic.
// k_gather spends ~130 instructions on every candidate; in the first doubling rounds of a
// repetitive text 75-99 % of all text positions are inert, and this kernel settles them with ~15.
// A block stages the candidates of its 8192 positions in shared memory and reserves their place
// in the list with one atomic (the list stays in text order up to the order of the blocks).
// ------------------------------------------------------------------------------------
struct PrefilterArgs {
  const u32 *rank;
  u32 n;
  u64 h;
  u32 round;
  u32 tiny_max;
  const u64 *state;
  const u32 *verdicts;
  const u64 *rho;
  u64 *rep;
  const u32 *hlist;   // labels of this round's huge groups (after k_huge_prepare)
  const u32 *hcount;
  u32 *cand;         // out: suffixes that need k_gather's full treatment
  u32 *cand_count;   // zeroed before launch
  u32 *counter;      // k_gather's counter word: inert suffixes are added to the live count ([0])
};

// The per-group look-ups (rho*, verdict) are what bounds a walk in text order: neighbouring suffixes belong to
// different groups, so a warp's 32 look-ups hit 32 different cache lines of tables that are indexed by label --
// one L1 wavefront per suffix (ncu: k_gather runs at 81 % of the L1 pipe and 23 % of DRAM).  Every block
// therefore starts by hashing this round's huge groups into a direct-mapped table in SHARED memory
// (label -> rho* | verdict bit): 32 look-ups then cost a few bank conflicts.  A group that loses its slot
// to another one is looked up in the global tables as before.
constexpr int PF_THREADS = 256, PF_WARPS = PF_THREADS / 32;
constexpr int PF_ITERS = 4;                        // a warp settles 128 * PF_ITERS consecutive suffixes per reservation
constexpr int PF_WCHUNK = 128 * PF_ITERS;          // ... and stages at most that many candidates
constexpr u32 PF_CACHE = 4096, PF_MAX_GROUPS = 2048;  // groups beyond that: no pre-filter (the table would thrash)
constexpr int PF_BLOCKS_PER_SM = 4;
constexpr u32 PF_VERDICT = 0x80000000u, PF_NO_RHO = 0x7fffffffu;  // values no label can have
constexpr size_t PF_SMEM = (size_t)PF_CACHE * 8 + (size_t)PF_WARPS * PF_WCHUNK * 4;

__device__ __forceinline__ u32 pf_slot(u32 lab) { return ((lab >> 8) * 2654435761u) >> 20; }   // 12 bits
__device__ __forceinline__ u32 pf_slot2(u32 lab) { return ((lab >> 8) * 0x85ebca6bu + 0x3c6ef372u) >> 20; }  // second choice
// entry of `lab`, or one with another tag if the group is not cached (two possible slots: at 1000 groups in 4096
// slots one in nine would lose a single slot to another group, and all of its members would go the slow way)
__device__ __forceinline__ unsigned long long pf_lookup(const unsigned long long *s_cache, u32 lab) {
  const unsigned long long e = s_cache[pf_slot(lab)];
  return ((u32)(e >> 32) == lab) ? e : s_cache[pf_slot2(lab)];
}

__device__ __noinline__ void pf_volunteer(u64 *rep, u32 lab, u32 round, u32 sfx) {
  const u32 hsh = mix32(sfx ^ (round * 0x9e3779b9u));
  atomicMin(reinterpret_cast<unsigned long long *>(rep + (lab / HUGE_M) * HUGE_REPS + ((hsh >> 8) & (HUGE_REPS - 1u))),
            (unsigned long long)rep_key(round, hsh >> 11, sfx));
}

// Warps work on their own: no block barrier inside the walk, one output reservation per warp and 512 suffixes
// (none at all when everything was inert), the inert count kept in a register until the end.
// ANYV: some huge group got a verdict this round (then a partner label that is a live huge label must be
// looked up as well: it counts only if its group got none).
template <bool ALIGNED, bool ANYV>
__device__ __forceinline__ void pf_walk(const PrefilterArgs &a, const unsigned long long *s_cache, u32 *s_stage) {
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const u32 lt = lanemask_lt();
  u32 *stage = s_stage + warp * PF_WCHUNK;
  const u32 dead_mask = RANK_DEAD | (a.tiny_max ? 1u : 0u);
  const u32 vol = a.round & 0xffu;
  const u32 nwc = (a.n + PF_WCHUNK - 1) / PF_WCHUNK;
  const u32 wstride = gridDim.x * PF_WARPS;
  u32 inert = 0;
  for (u32 wc = blockIdx.x * PF_WARPS + warp; wc < nwc; wc += wstride) {
    uint4 wv[PF_ITERS], rv[PF_ITERS];
#pragma unroll
    for (int it = 0; it < PF_ITERS; ++it) {
      const u32 i0 = wc * (u32)PF_WCHUNK + (u32)it * 128u + lane * 4u;
      wv[it] = make_uint4(RANK_DEAD, RANK_DEAD, RANK_DEAD, RANK_DEAD);
      rv[it] = make_uint4(0, 0, 0, 0);
      if (i0 + 4u <= a.n) {
        wv[it] = ld_stream_u128(a.rank + i0);
      } else {
        u32 t[4] = {RANK_DEAD, RANK_DEAD, RANK_DEAD, RANK_DEAD};
        for (int j = 0; j < 4; ++j) if (i0 + j < a.n) t[j] = a.rank[i0 + j];
        wv[it] = make_uint4(t[0], t[1], t[2], t[3]);
      }
      const u64 t0 = (u64)i0 + a.h;
      if (ALIGNED && t0 + 4u <= a.n) {
        rv[it] = ld_stream_u128(a.rank + t0);
      } else {
        u32 t[4] = {0, 0, 0, 0};
        for (int j = 0; j < 4; ++j) if (t0 + j < a.n) t[j] = __ldg(a.rank + t0 + j);
        rv[it] = make_uint4(t[0], t[1], t[2], t[3]);
      }
    }
    u32 nst = 0;  // candidates staged by this warp so far
#pragma unroll
    for (int it = 0; it < PF_ITERS; ++it) {
      const u32 i0 = wc * (u32)PF_WCHUNK + (u32)it * 128u + lane * 4u;
      const u32 w[4] = {wv[it].x, wv[it].y, wv[it].z, wv[it].w};
      const u32 r2[4] = {rv[it].x, rv[it].y, rv[it].z, rv[it].w};
      u32 slow = 0;  // bit j: suffix i0 + j goes to k_gather
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool live = !(w[j] & dead_mask) && w[j] != 0u;
        const unsigned long long ent = pf_lookup(s_cache, w[j]);
        bool is_inert = live && (u32)(ent >> 32) == w[j] && (u32)ent == (r2[j] & RANK_MASK);  // (a cached label is a huge one)
        if (ANYV) {
          if (is_inert && needs_state(r2[j])) {
            const unsigned long long e2 = pf_lookup(s_cache, r2[j]);
            is_inert = (u32)(e2 >> 32) == r2[j] && !((u32)e2 & PF_VERDICT);  // not cached: let k_gather look it up
          }
        }
        if (is_inert) {
          ++inert;
          const u32 sfx = i0 + j;
          if (((sfx * 0x9e3779b1u) >> 24) == vol) pf_volunteer(a.rep, w[j], a.round, sfx);  // same volunteers as k_gather
        } else if (live) {
          slow |= 1u << j;
        }
      }
      // stage the slow ones: (lane, j) order = text order inside the warp's 128 positions
      if (__any_sync(0xffffffffu, slow != 0u)) {
        u32 before = 0, total = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const u32 b = __ballot_sync(0xffffffffu, (slow >> j) & 1u);
          before += (u32)__popc(b & lt);
          total += (u32)__popc(b);
        }
        u32 o = nst + before;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if ((slow >> j) & 1u) stage[o++] = i0 + j;
        nst += total;
      }
    }
    if (nst) {  // warp-uniform
      u32 base = 0;
      if (lane == 0) base = atomicAdd(a.cand_count, nst);
      base = __shfl_sync(0xffffffffu, base, 0);
      __syncwarp();
      for (u32 x = lane; x < nst; x += 32u) a.cand[base + x] = stage[x];
      __syncwarp();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) inert += __shfl_xor_sync(0xffffffffu, inert, o);
  if (lane == 0 && inert) atomicAdd(reinterpret_cast<unsigned long long *>(a.counter), (unsigned long long)inert);  // live count, low word
}

template <bool ALIGNED>
__global__ void __launch_bounds__(PF_THREADS, PF_BLOCKS_PER_SM) k_prefilter(const PrefilterArgs a) {
  extern __shared__ __align__(16) unsigned char pf_smem[];
  unsigned long long *s_cache = reinterpret_cast<unsigned long long *>(pf_smem);  // label << 32 | (rho* or PF_NO_RHO) | PF_VERDICT
  u32 *s_stage = reinterpret_cast<u32 *>(s_cache + PF_CACHE);                       // [PF_WARPS][PF_WCHUNK]
  const u32 tid = threadIdx.x;
  const bool anyv = __ldg(a.verdicts) != 0u;
  for (u32 i = tid; i < PF_CACHE; i += PF_THREADS) s_cache[i] = 0ull;
  __syncthreads();
  {
    const u32 hc = __ldg(a.hcount);
    for (u32 j = tid; j < hc; j += PF_THREADS) {
      const u32 lab = __ldg(a.hlist + j);
      const u64 e = __ldg(a.rho + (lab / HUGE_M));
      const bool verdict = anyv && (u32)(__ldg(a.state + (lab / HUGE_M)) >> 32) == a.round;
      u32 val = ((u32)(e >> 32) == a.round) ? (u32)e : PF_NO_RHO;
      if (verdict) val = PF_VERDICT | PF_NO_RHO;
      const unsigned long long entry = ((unsigned long long)lab << 32) | val;
      if (atomicCAS(&s_cache[pf_slot(lab)], 0ull, entry) != 0ull) atomicCAS(&s_cache[pf_slot2(lab)], 0ull, entry);  // first come, first served
    }
  }
  __syncthreads();
  if (anyv) pf_walk<ALIGNED, true>(a, s_cache, s_stage);
  else pf_walk<ALIGNED, false>(a, s_cache, s_stage);
}

// Sparse mode, after k_gather: fill in the missing second key halves, then histogram.
__global__ void __launch_bounds__(256) k_lazy_fill(const KeyGen g, const i32 *__restrict__ sa0, const u32 *__restrict__ todo,
                                                   const u32 *__restrict__ todo_count, u64 *__restrict__ keys) {
  const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= *todo_count) return;
  keys[todo[2 * q]] |= (u64)lazy_label(g, sa0, todo[2 * q + 1]);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_hist_keys(const u64 *__restrict__ keys, u32 L, int npass, u32 *__restrict__ ghist) {
  __shared__ u32 shist[MAX_PASSES * RADIX];
  for (int i = threadIdx.x; i < npass * RADIX; i += THREADS) shist[i] = 0;
  __syncthreads();
  const u32 stride = gridDim.x * THREADS;
  const u32 iters = (L + stride - 1) / stride;
  u32 l = blockIdx.x * THREADS + threadIdx.x;
  for (u32 it = 0; it < iters; ++it, l += stride) {
    const bool valid = l < L;
    hist_add(shist, valid ? keys[l] : 0ull, valid, npass);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * RADIX; i += THREADS)
    if (shist[i]) atomicAdd(&ghist[i], shist[i]);
}

// ------------------------------------------------------------------------------------
// Step 2b (after the sort): SA slot of every sorted element.
// The sorted list holds, group by group (a RUN = maximal stretch with the same label), every
// live member except the inert ones.  With G[label] = (gs, ge) and, for a huge group, its rho*:
//   second key half < rho* (or no rho*): slot = gs + (index inside the run)
//   second key half > rho*             : slot = ge - (elements behind it in the run)
// Index inside the run needs the run head (forward max-scan), elements behind it the run tail
// (backward min-scan); k_run_summary + k_tail_scan prepare both carries across tiles ("last run
// head before tile t", "first run tail after tile t"), so no block waits for another one.  The two ends of the inert block move inwards past the sorted members:
// the last "<" element and the first ">" element of a run record the new ends in a small
// update list, applied by k_apply_g after this kernel (G is read here, so it is not written).
// Bit 31 of the slot word tells the rebuild that the run has an inert block.
// ------------------------------------------------------------------------------------
struct SlotArgs {
  const u64 *keys;
  u32 S;
  u32 lab_bits;
  const u64 *G;
  const u64 *rho;
  u32 round;
  int filter;
  u32 *slots;
  u32 *tile_rtail;      // [tiles] k_run_summary: first run-tail index inside the tile
  u32 *tile_rhead;      // [tiles] k_run_summary: last run-head index + 1 inside the tile (or 0)
  const u32 *next_rtail;  // [tiles] k_tail_scan: first run-tail index in any later tile
  const u32 *prev_rhead;  // [tiles] k_tail_scan: last run-head index + 1 in any earlier tile
  u32 *gupd;            // triples (label, 0 = first slot / 1 = last slot, value)
  u32 *gupd_count;
};
constexpr u32 SLOT_HAS_RHO = 0x80000000u;

template <int THREADS, int IPT>
__global__ void __launch_bounds__(THREADS) k_run_summary(const SlotArgs a) {
  constexpr int WARPS = THREADS / 32;
  __shared__ u32 s_w[WARPS], s_h[WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 l0 = blockIdx.x * (u32)(THREADS * IPT) + (u32)tid * IPT;
  u32 v = NO_TAIL_IDX, hd = 0;  // first run-tail index, last run-head index + 1
  if (l0 < a.S) {
    const u32 nvalid = min((u32)IPT, a.S - l0);
    u32 prev = (l0 > 0) ? (u32)(a.keys[l0 - 1] >> a.lab_bits) : 0u;
    u32 lab = (u32)(a.keys[l0] >> a.lab_bits);
    for (u32 j = 0; j < nvalid; ++j) {
      const u32 l = l0 + j;
      const bool last = l + 1 >= a.S;
      const u32 nlab = last ? 0u : (u32)(a.keys[l + 1] >> a.lab_bits);
      if (l == 0 || lab != prev) hd = l + 1u;
      if ((last || nlab != lab) && v == NO_TAIL_IDX) v = l;
      prev = lab;
      lab = nlab;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    hd = max(hd, __shfl_xor_sync(0xffffffffu, hd, o));
  }
  if (lane == 0) { s_w[warp] = v; s_h[warp] = hd; }
  __syncthreads();
  if (tid == 0) {
    u32 m = NO_TAIL_IDX, h = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { m = min(m, s_w[w]); h = max(h, s_h[w]); }
    a.tile_rtail[blockIdx.x] = m;
    a.tile_rhead[blockIdx.x] = h;
  }
}

template <int THREADS, int IPT>
__global__ void __launch_bounds__(THREADS, 2) k_slots(const SlotArgs a) {
  constexpr int WARPS = THREADS / 32;
  constexpr int TILE = THREADS * IPT;
  __shared__ u32 s_wh[WARPS], s_wt[WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 tile = blockIdx.x;
  const u32 S = a.S;
  const u32 l0 = tile * (u32)TILE + (u32)tid * IPT;
  u32 nvalid = 0;
  if (l0 < S) nvalid = min((u32)IPT, S - l0);
  // labels (first key half) of the IPT elements and of one neighbour on each side; second halves
  u32 lab[IPT + 2], r2[IPT + 2];
  const u64 r2mask = (1ull << a.lab_bits) - 1ull;
#pragma unroll
  for (int j = 0; j < IPT + 2; ++j) {
    const i64 l = (i64)l0 + j - 1;
    u64 kx = 0;
    if (l >= 0 && l < (i64)S && (j == 0 ? nvalid > 0 : (u32)(j - 1) <= nvalid)) kx = a.keys[l];
    lab[j] = (u32)(kx >> a.lab_bits);
    r2[j] = (u32)(kx & r2mask);
  }
  // run flags: bit j = element l0+j starts a run (beyond-the-end counts as a start)
  u32 f = 0;
#pragma unroll
  for (int j = 0; j <= IPT; ++j) {
    const u32 l = l0 + j;
    f |= ((l == 0 || l >= S || lab[j + 1] != lab[j]) ? 1u : 0u) << j;
  }
  u32 th = 0, tt = NO_TAIL_IDX;  // last run head index + 1, first run tail index
#pragma unroll
  for (int j = 0; j < IPT; ++j)
    if ((u32)j < nvalid && ((f >> j) & 1u)) th = l0 + j + 1u;
#pragma unroll
  for (int j = IPT - 1; j >= 0; --j)
    if ((u32)j < nvalid && ((f >> (j + 1)) & 1u)) tt = l0 + j;
  u32 ih = th, it = tt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 yh = __shfl_up_sync(0xffffffffu, ih, o);
    const u32 yt = __shfl_down_sync(0xffffffffu, it, o);
    if (lane >= o) ih = max(ih, yh);
    if (lane + o < 32) it = min(it, yt);
  }
  if (lane == 31) s_wh[warp] = ih;
  if (lane == 0) s_wt[warp] = it;
  u32 eh = __shfl_up_sync(0xffffffffu, ih, 1);
  u32 et = __shfl_down_sync(0xffffffffu, it, 1);
  if (lane == 0) eh = 0;
  if (lane == 31) et = NO_TAIL_IDX;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < warp) eh = max(eh, s_wh[w]);
    if (w > warp) et = min(et, s_wt[w]);
  }
  // carries across tiles come from k_run_summary + k_tail_scan: no block waits for another one
  et = min(et, a.next_rtail[tile]);
  eh = max(eh, a.prev_rhead[tile]);
  u32 rt[IPT];  // run tail index of every element
  {
    u32 cur = et;
#pragma unroll
    for (int j = IPT - 1; j >= 0; --j) {
      if ((u32)j < nvalid && ((f >> (j + 1)) & 1u)) cur = l0 + j;
      rt[j] = cur;
    }
  }
  u32 rh = eh;  // run head index + 1
  u32 gs = 0, ge = 0, rho = 0;
  bool has_rho = false;
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    if ((u32)j < nvalid) {
      const u32 l = l0 + j;
      const bool start = (f >> j) & 1u;
      if (start) rh = l + 1u;
      if (start || j == 0) {  // (re)load the group's table entry
        const u32 lb = lab[j + 1];
        const u64 g = __ldcg(a.G + lb);
        gs = (u32)g; ge = (u32)(g >> 32);
        has_rho = false;
        if (a.filter && is_huge_label(lb)) {
          const u64 e = __ldcg(a.rho + (lb / HUGE_M));
          if ((u32)(e >> 32) == a.round) { has_rho = true; rho = (u32)e; }
        }
      }
      const u32 A = l - (rh - 1u), B = rt[j] - l;
      const bool less = !has_rho || r2[j + 1] < rho;
      const u32 slot = less ? (gs + A) : (ge - B);
      a.slots[l] = slot | (has_rho ? SLOT_HAS_RHO : 0u);
      if (has_rho) {
        const bool is_tail = (f >> (j + 1)) & 1u;
        if (less && (is_tail || !(r2[j + 2] < rho))) {   // last element left of the inert block
          const u32 q = atomicAdd(a.gupd_count, 1u);
          a.gupd[3 * q] = lab[j + 1]; a.gupd[3 * q + 1] = 0u; a.gupd[3 * q + 2] = slot + 1u;
        }
        if (!less && (start || r2[j] < rho)) {           // first element right of it
          const u32 q = atomicAdd(a.gupd_count, 1u);
          a.gupd[3 * q] = lab[j + 1]; a.gupd[3 * q + 1] = 1u; a.gupd[3 * q + 2] = slot - 1u;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_apply_g(const u32 *__restrict__ gupd, const u32 *__restrict__ count, u64 *__restrict__ G) {
  const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= *count) return;
  u32 *half = reinterpret_cast<u32 *>(G + gupd[3 * q]);
  half[gupd[3 * q + 1]] = gupd[3 * q + 2];
}

// ------------------------------------------------------------------------------------
// Step 3 (after the sort): re-flag, re-label, finalise, compact.  Sorted live elements l
// (key[l], suffix[l]) sit in SA slots pos[l] (ascending).  With
//   flag[l] = key[l] != key[l-1]           (round 0: also around short suffixes)
//   head[l] = slot of the nearest flagged element at or before l      (forward max-scan)
//   tail[l] = slot of the nearest element at or after l whose successor is flagged
//                                                                      (backward min-scan)
// the new group of l occupies slots [head, tail]:
//   head == tail       -> unique: SA[head] = suffix, rank[suffix] = DEAD | head+1
//   old label in range -> nothing to write (the group keeps its label)
//   otherwise          -> rank[suffix] = middle of the range + 1
// and the slots of the non-unique elements are compacted for the next round.
// Neither scan waits on another block: a first light kernel (k_tail_summary) records every
// tile's last head slot and first tail slot, and a one-block scan (k_tail_scan) turns them into
// "last head slot before tile t" and "first tail slot after tile t".
// ------------------------------------------------------------------------------------
struct RebuildArgs {
  const u64 *keys;
  const u32 *sufx;
  const u32 *pos_in;   // k_slots output (bit 31: the run has an inert block); null in round 0 (slot == index)
  u32 L;
  u32 short_from;      // round 0: suffix indices >= short_from are short (forced singletons)
  u32 lab_bits;        // old label = key >> lab_bits (rounds >= 1)
  u32 *rank;
  i32 *SA;
  u64 *G;              // [n + 2] slot range of every live group, indexed by its label
  u32 *hlist;          // labels of the huge groups created in this round are appended here
  u32 *hcount;
  u64 *rep;            // [n / HUGE_M + 2][HUGE_REPS] their first member becomes the representative
  u32 round;           // round this rebuild belongs to (0 = round 0)
  HugeKeyTable hkt;    // round 0: huge groups register their key here (k_rank_huge0 labels the members)
  u32 tiny_max;        // groups of at most this many suffixes move to the bag (0: no bag)
  u64 *bag_desc;       // (first slot | size << 32) of every tiny group formed by this rebuild
  u32 *bag_desc_count;
  u32 *survivors;      // [1] k_tail_summary: number of elements that stay live after this round
  u32 *tile_tail;      // [tiles] k_tail_summary: first tail slot inside the tile (or NO_TAIL)
  u32 *tile_head;      // [tiles] k_tail_summary: last head slot + 1 inside the tile (or 0)
  const u32 *next_tail;  // [tiles] k_tail_scan: first tail slot in any later tile
  const u32 *prev_head;  // [tiles] k_tail_scan: last head slot + 1 in any earlier tile
  bool sa_holds_sufx;  // round 0: sufx IS the SA (slot l already holds suffix sufx[l]): no SA writes at all
  u32 *surv_list;      // RB_SPARSE: the suffixes that are not unique yet are listed here (any order) ...
  u32 *surv_count;     // ... so that round 1 walks them instead of all n text positions
  // round 0: k_tail_summary also lists the INDICES of the elements that are not unique (while they are rare); when round 0
  // turns out sparse and all groups are small, k_rebuild_list handles just those instead of a pass over all n elements
  u32 *live_idx;       // [<= L / 4] indices into the sorted sequence, any order
  u32 *live_idx_ctl;   // [0] entries, [1] != 0: some warp saw too many to list, [2] != 0: k_rebuild_list met a large group
};
constexpr u32 LIVE_IDX_PER_WARP = 64;  // a warp (32 * IPT elements) lists at most this many; more = not a sparse text
constexpr u32 LIST_GROUP_MAX = 64;     // k_rebuild_list walks groups up to this size, larger ones go the general way

// Loads the IPT consecutive elements of this thread plus one neighbour on each side and
// returns the flag bits f (bit j = flag of element l0 + j, j = 0..IPT; beyond-the-end counts
// as flagged).  kx[j+1] / sx[j+1] belong to element l0 + j.
template <int IPT, bool ROUND0>
__device__ __forceinline__ u32 load_and_flag(const RebuildArgs &a, u32 l0, u64 (&kx)[IPT + 2], u32 (&sx)[IPT + 2]) {
  const u32 L = a.L;
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
  } else {
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const bool ok = l0 + j < L;
      kx[j + 1] = ok ? a.keys[l0 + j] : 0;
      sx[j + 1] = ok ? a.sufx[l0 + j] : 0;
    }
  }
  const bool has_prev = (l0 > 0) && (l0 <= L);
  const bool has_next = (l0 + IPT < L);
  kx[0] = has_prev ? a.keys[l0 - 1] : 0;
  sx[0] = has_prev ? a.sufx[l0 - 1] : 0;
  kx[IPT + 1] = has_next ? a.keys[l0 + IPT] : 0;
  sx[IPT + 1] = has_next ? a.sufx[l0 + IPT] : 0;
  u32 f = 0;
#pragma unroll
  for (int j = 0; j <= IPT; ++j) {
    const u32 l = l0 + j;
    bool fl = (l == 0) || (l >= L) || (kx[j + 1] != kx[j]);
    if (ROUND0) fl = fl || (sx[j + 1] >= a.short_from) || (sx[j] >= a.short_from);
    f |= (fl ? 1u : 0u) << j;
  }
  return f;
}

template <int THREADS, int IPT, bool ROUND0>
__global__ void __launch_bounds__(THREADS) k_tail_summary(const RebuildArgs a) {
  constexpr int WARPS = THREADS / 32;
  __shared__ u32 s_w[WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 tile = blockIdx.x;
  const u32 l0 = tile * (u32)(THREADS * IPT) + (u32)tid * IPT;
  u64 kx[IPT + 2];
  u32 sx[IPT + 2];
  const u32 f = load_and_flag<IPT, ROUND0>(a, l0, kx, sx);
  u32 v = NO_TAIL, surv = 0, hd = 0;  // first tail slot, survivors, last head slot + 1
  if (l0 < a.L) {
    const u32 nvalid = min((u32)IPT, a.L - l0);
    const u32 vm = (1u << nvalid) - 1u;
    const u32 tails = (f >> 1) & vm;  // element j is a tail iff element j+1 is flagged
    if (tails) {
      const u32 l = l0 + (u32)(__ffs(tails) - 1);
      v = ROUND0 ? l : (a.pos_in[l] & ~SLOT_HAS_RHO);
    }
    const u32 heads = f & vm;
    if (heads) {
      const u32 l = l0 + (31u - (u32)__clz(heads));
      hd = (ROUND0 ? l : (a.pos_in[l] & ~SLOT_HAS_RHO)) + 1u;
    }
    surv = (u32)__popc(~(f & (f >> 1)) & vm);  // not (head and tail) = not unique yet
  }
  const u32 my_surv = surv;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    hd = max(hd, __shfl_xor_sync(0xffffffffu, hd, o));
    surv += __shfl_xor_sync(0xffffffffu, surv, o);
  }
  bool too_many = false;  // (warp-uniform) this warp saw more non-unique elements than it lists
  if (ROUND0 && a.live_idx != nullptr && surv != 0u) {  // (warp-uniform) list the warp's non-unique elements while they are few
    if (surv > LIVE_IDX_PER_WARP) {
      too_many = true;
    } else {
      u32 inc = my_surv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      u32 base = 0;
      if (lane == 31) base = atomicAdd(a.live_idx_ctl, inc);
      base = __shfl_sync(0xffffffffu, base, 31) + inc - my_surv;
      if (my_surv) {
        const u32 nvalid = min((u32)IPT, a.L - l0);
        u32 m = ~(f & (f >> 1)) & ((1u << nvalid) - 1u);
        while (m) {
          const u32 j = (u32)__ffs(m) - 1u;
          m &= m - 1u;
          a.live_idx[base++] = l0 + j;
        }
      }
    }
  }
  __shared__ u32 s_s[WARPS], s_h[WARPS], s_o[WARPS];
  if (lane == 0) { s_w[warp] = v; s_s[warp] = surv; s_h[warp] = hd; s_o[warp] = too_many ? 1u : 0u; }
  __syncthreads();
  if (tid == 0) {
    u32 m = NO_TAIL, t = 0, h = 0, o = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { m = min(m, s_w[w]); t += s_s[w]; h = max(h, s_h[w]); o |= s_o[w]; }
    a.tile_tail[tile] = m;
    a.tile_head[tile] = h;
    if (t) atomicAdd(a.survivors, t);
    // one look per block, a store only while the word is still zero: every warp of a repetitive text ends up here, and four
    // million stores to one address cost round 0 of rep_1G 4 ms
    if (ROUND0 && o && __ldcg(a.live_idx_ctl + 1) == 0u) a.live_idx_ctl[1] = 1u;
  }
}

// next_tail[t] = min over tiles t' > t of tile_tail[t'] (slots ascend, so the minimum is the nearest);
// if tile_head != null also prev_head[t] = max over tiles t' < t of tile_head[t'].
// With these two per-tile carries the rebuild kernels need no look-back between their blocks.
__global__ void __launch_bounds__(1024) k_tail_scan(const u32 *__restrict__ tile_tail, u32 *__restrict__ next_tail, u32 tiles,
                                                    const u32 *__restrict__ tile_head, u32 *__restrict__ prev_head) {
  // One block; warp w owns a contiguous span of the tiles and walks it in rows of 32 (coalesced loads, a shuffle scan per
  // row, a running carry), after a first walk that gives every warp the carry it starts from.  (One thread per chunk with
  // sequential dependent loads took 76-120 us per launch at 1 GiB -- 25 launches per build.)
  __shared__ u32 s_max[32], s_min[32];
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const u32 span = (((tiles + 31u) / 32u) + 31u) / 32u * 32u;  // tiles per warp, a multiple of 32
  const u32 lo = min(tiles, warp * span), hi = min(tiles, lo + span);
  u32 m = 0, mn = NO_TAIL;
  for (u32 i = lo + lane; i < hi; i += 32u) {
    if (tile_head != nullptr) m = max(m, tile_head[i]);
    mn = min(mn, tile_tail[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) { s_max[warp] = m; s_min[warp] = mn; }
  __syncthreads();
  u32 cmax = 0, cmin = NO_TAIL;  // max over the warps in front of mine, min over the warps behind it
  for (u32 w = 0; w < 32u; ++w) {
    if (w < warp) cmax = max(cmax, s_max[w]);
    if (w > warp) cmin = min(cmin, s_min[w]);
  }
  if (tile_head != nullptr) {  // prev_head[t] = max over tiles t' < t of tile_head[t']
    u32 carry = cmax;
    for (u32 base = lo; base < hi; base += 32u) {
      const u32 i = base + lane;
      u32 inc = (i < hi) ? tile_head[i] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (u32)o) inc = max(inc, y);
      }
      u32 exc = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) exc = 0u;
      if (i < hi) prev_head[i] = max(carry, exc);
      carry = max(carry, __shfl_sync(0xffffffffu, inc, 31));
    }
  }
  {  // next_tail[t] = min over tiles t' > t of tile_tail[t'] (slots ascend, so the minimum is the nearest)
    u32 carry = cmin;
    const u32 rows = (hi - lo + 31u) / 32u;
    for (u32 r = rows; r > 0; --r) {
      const u32 i = lo + (r - 1u) * 32u + lane;
      u32 inc = (i < hi) ? tile_tail[i] : NO_TAIL;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + (u32)o < 32u) inc = min(inc, y);
      }
      u32 exc = __shfl_down_sync(0xffffffffu, inc, 1);
      if (lane == 31) exc = NO_TAIL;
      if (i < hi) next_tail[i] = min(carry, exc);
      carry = min(carry, __shfl_sync(0xffffffffu, inc, 0));
    }
  }
}

// Round 0 of a sparse text (few elements are not unique, all of them in small groups): the rebuild proper for just the
// listed elements -- what k_rebuild<ROUND0, RB_SPARSE> does for them, without its pass over all n elements (which reads
// 12 bytes per element to find that nearly every one is unique: 1.2 of rand_256M's 11.9 ms).  One thread per listed
// element walks to the two ends of its group (flags as in load_and_flag); a group of more than LIST_GROUP_MAX elements
// raises ctl[2] and the host runs the general kernel instead (every write here is one the general kernel repeats).
__global__ void __launch_bounds__(256) k_rebuild_list(const RebuildArgs a, const u32 count) {
  const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  const u32 l = a.live_idx[q];
  const u32 L = a.L;
  auto flagged = [&](u32 x) -> bool {  // element x starts a group (x == L: beyond the end)
    if (x == 0u || x >= L) return true;
    return a.keys[x] != a.keys[x - 1u] || a.sufx[x] >= a.short_from || a.sufx[x - 1u] >= a.short_from;
  };
  u32 h = l, t = l;
  u32 steps = 0;
  while (!flagged(h)) { --h; if (++steps > LIST_GROUP_MAX) break; }
  while (steps <= LIST_GROUP_MAX && !flagged(t + 1u)) { ++t; if (++steps > LIST_GROUP_MAX) break; }
  if (steps > LIST_GROUP_MAX || t - h + 1u >= HUGE_T) {
    a.live_idx_ctl[2] = 1u;
    return;
  }
  // group = slots [h, t] (round 0: slot == index), t > h because the element is not unique
  const u32 lab = pick_label(h, t, 0u, 0u);
  const u32 sfx = a.sufx[l];
  a.rank[sfx] = lab;
  a.surv_list[atomicAdd(a.surv_count, 1u)] = sfx;
  if (l == h) a.G[lab] = (u64)h | ((u64)t << 32);
}

// MODE (the survivor count of the round is known from k_tail_summary before the launch):
//   RB_NORMAL  as described above
//   RB_FINAL   nothing survives: nobody will read a rank again, only SA is written
//   RB_SPARSE  round 0 with few survivors: SA gets every element (a complete round-0 order),
//              rank[] (pre-zeroed) only the labels of the survivors; labels of the unique
//              suffixes are recomputed on demand (lazy_label) instead of being scattered
enum { RB_NORMAL = 0, RB_FINAL = 1, RB_SPARSE = 2 };
template <int THREADS, int IPT, bool ROUND0, int MODE>
__global__ void __launch_bounds__(THREADS, 2) k_rebuild(const RebuildArgs a) {
  constexpr int WARPS = THREADS / 32;
  constexpr int TILE = THREADS * IPT;
  __shared__ u32 s_wh[WARPS], s_wt[WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 tile = blockIdx.x;
  const u32 L = a.L;
  const u32 l0 = tile * (u32)TILE + (u32)tid * IPT;

  u64 kx[IPT + 2];
  u32 sx[IPT + 2];
  u32 px[IPT];
  const u32 f = load_and_flag<IPT, ROUND0>(a, l0, kx, sx);
  if (ROUND0) {
#pragma unroll
    for (int j = 0; j < IPT; ++j) px[j] = l0 + j;
  } else if (l0 + IPT <= L) {
#pragma unroll
    for (int j = 0; j < IPT; j += 4) {
      const uint4 v = *reinterpret_cast<const uint4 *>(a.pos_in + l0 + j);
      px[j] = v.x; px[j + 1] = v.y; px[j + 2] = v.z; px[j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < IPT; ++j) px[j] = (l0 + j < L) ? a.pos_in[l0 + j] : 0;
  }
  u32 hr = 0;  // bit j: element j belongs to a run with an inert block (its old label stays with that block)
  if (!ROUND0) {
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      hr |= (px[j] >> 31) << j;
      px[j] &= ~SLOT_HAS_RHO;
    }
  }
  u32 nvalid = 0;
  if (l0 < L) nvalid = min((u32)IPT, L - l0);

  // ---- thread aggregates ------------------------------------------------------------------
  u32 th = 0;        // last flagged slot + 1
  u32 tt = NO_TAIL;  // first tail slot
#pragma unroll
  for (int j = IPT - 1; j >= 0; --j)
    if ((u32)j < nvalid && ((f >> (j + 1)) & 1u)) tt = px[j];
#pragma unroll
  for (int j = 0; j < IPT; ++j)
    if ((u32)j < nvalid && ((f >> j) & 1u)) th = px[j] + 1u;
  // ---- block scans: forward max (head) and backward min (tail); the carries across tiles were
  // prepared by k_tail_summary + k_tail_scan, so no block waits for another one --------------------
  u32 ih = th, it = tt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 yh = __shfl_up_sync(0xffffffffu, ih, o);
    const u32 yt = __shfl_down_sync(0xffffffffu, it, o);
    if (lane >= o) ih = max(ih, yh);
    if (lane + o < 32) it = min(it, yt);
  }
  if (lane == 31) s_wh[warp] = ih;
  if (lane == 0) s_wt[warp] = it;
  u32 eh = __shfl_up_sync(0xffffffffu, ih, 1);
  u32 et = __shfl_down_sync(0xffffffffu, it, 1);
  if (lane == 0) eh = 0;
  if (lane == 31) et = NO_TAIL;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < warp) eh = max(eh, s_wh[w]);
    if (w > warp) et = min(et, s_wt[w]);
  }
  eh = max(eh, a.prev_head[tile]);  // head slot + 1 of the group that reaches into this thread's elements
  et = min(et, a.next_tail[tile]);  // first tail slot after this thread's elements

  // ---- emit ------------------------------------------------------------------------------------
  u32 tl[IPT];  // tail slot of every element
  {
    u32 cur = et;
#pragma unroll
    for (int j = IPT - 1; j >= 0; --j) {
      if ((u32)j < nvalid && ((f >> (j + 1)) & 1u)) cur = px[j];
      tl[j] = cur;
    }
  }
  // tiny groups headed here: one descriptor each, reserved with one atomic per warp (per block: no faster)
  u32 dbase = 0;
  if (a.tiny_max) {
    u32 nd = 0, head = eh;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      if ((u32)j < nvalid && ((f >> j) & 1u)) {
        head = px[j] + 1u;
        const u32 size = tl[j] + 2u - head;
        nd += (size > 1u && size <= a.tiny_max) ? 1u : 0u;
      }
    }
    u32 inc = nd;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    u32 wb = 0;
    if (lane == 31 && inc) wb = atomicAdd(a.bag_desc_count, inc);
    dbase = __shfl_sync(0xffffffffu, wb, 31) + inc - nd;
  }
  // sparse round 0: the few survivors are listed (one reservation per warp), round 1 then walks that list
  u32 sbase = 0;
  if (MODE == RB_SPARSE) {
    const u32 vm = (nvalid >= 32u) ? 0xffffffffu : ((1u << nvalid) - 1u);
    const u32 ns = (u32)__popc(~(f & (f >> 1)) & vm);  // not (head and tail) = not unique yet
    u32 inc = ns;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    u32 wb = 0;
    if (lane == 31 && inc) wb = atomicAdd(a.surv_count, inc);
    sbase = __shfl_sync(0xffffffffu, wb, 31) + inc - ns;
  }
  u32 head = eh;  // head slot + 1
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    if ((u32)j < nvalid) {
      if ((f >> j) & 1u) head = px[j] + 1u;
      const u32 s1 = head, e1 = tl[j] + 1u;  // label range of the group: [s1, e1]
      if (MODE == RB_SPARSE && s1 != e1) a.surv_list[sbase++] = sx[j + 1];
      const bool write_sa = !(ROUND0 && a.sa_holds_sufx);
      if (s1 == e1) {
        if (write_sa) a.SA[px[j]] = (i32)sx[j + 1];
        if (MODE == RB_NORMAL) a.rank[sx[j + 1]] = RANK_DEAD | s1;
      } else {
        const u32 old = ROUND0 ? 0u : (u32)(kx[j + 1] >> a.lab_bits);
        const bool inert_owner = (hr >> j) & 1u;  // the old label stays with the run's inert block
        const u32 size = e1 - s1 + 1u;
        const bool tiny = size <= a.tiny_max;
        // a medium group keeps its label while the label stays inside its range; huge labels are owned
        // by inert blocks, tiny groups always carry the canonical label of their range
        const bool keep = !ROUND0 && !inert_owner && !tiny && size < HUGE_T && old >= s1 && old <= e1 &&
                          !is_huge_label(old) && (a.tiny_max == 0u || !(old & 1u));
        const u32 lab = keep ? old : pick_label(s1 - 1u, e1 - 1u, inert_owner ? old : 0u, a.tiny_max);
        const bool by_table = ROUND0 && is_huge_label(lab);  // members are labelled by k_rank_huge0
        if (lab != old && !by_table) a.rank[sx[j + 1]] = lab;
        if (tiny && write_sa) a.SA[px[j]] = (i32)sx[j + 1];  // provisional order: the bag is backed by SA
        if ((f >> j) & 1u) {  // group head: publish the group's slot range
          if (tiny) {
            a.bag_desc[dbase++] = (u64)(s1 - 1u) | ((u64)size << 32);
          } else {
            a.G[lab] = (u64)(s1 - 1u) | ((u64)(e1 - 1u) << 32);
            if (!keep && is_huge_label(lab)) {
              a.hlist[atomicAdd(a.hcount, 1u)] = lab;
              a.rep[(lab / HUGE_M) * HUGE_REPS] = rep_key(a.round + 1u, 0u, sx[j + 1]);  // the other slots are filled by inert volunteers
              if (ROUND0) hkt_insert(a.hkt, kx[j + 1], lab);
            }
          }
        }
        if (MODE == RB_SPARSE && write_sa) a.SA[px[j]] = (i32)sx[j + 1];
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// The bag (see pick_label): tiny groups, refined without the radix sort.
//   k_bag_append   descriptors written by a rebuild -> flattened (suffix, slot | HEAD) entries
//   k_bag_gather   r2 = label(suffix + h) for every entry (the one random read of the round)
//   k_bag_refine   one thread per entry, one block per BAG_TILE entries (+ TINY_MAX of overlap, so
//                  that a group never straddles a block): order the members of each group by r2 (rank
//                  by counting, groups have <= TINY_MAX members), split into runs of equal r2,
//                  finalise the unique ones, re-label and re-append the others.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bag_append(const u64 *__restrict__ desc, const u32 *__restrict__ desc_count,
                                                    const i32 *__restrict__ SA, u32 *__restrict__ bag_sufx,
                                                    u32 *__restrict__ bag_pos, u32 *__restrict__ bag_count) {
  // a warp expands 32 descriptors together: one reservation, then 32 entries per step, each lane
  // finding the descriptor of its entry by binary search over the warp's running sizes
  __shared__ u32 s_tot[8];
  __shared__ u32 s_base;
  const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const u32 nd = *desc_count;
  if (blockIdx.x * blockDim.x >= nd) return;  // whole block
  const u64 d = (q < nd) ? desc[q] : 0ull;
  const u32 s = (u32)d, size = (u32)(d >> 32);
  u32 inc = size;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)lane >= o) inc += y;
  }
  const u32 total = __shfl_sync(0xffffffffu, inc, 31);
  // one reservation per block (all blocks of the grid add to one counter)
  if (lane == 0) s_tot[warp] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_tot[w];
    s_base = t ? atomicAdd(bag_count, t) : 0u;
  }
  __syncthreads();
  u32 base = s_base;
  for (u32 w = 0; w < warp; ++w) base += s_tot[w];
  for (u32 e0 = 0; e0 < total; e0 += 32u) {
    const u32 e = e0 + lane;
    // first descriptor whose inclusive running size exceeds e
    u32 lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const u32 probe = __shfl_sync(0xffffffffu, inc, (int)(lo + step - 1u));
      if (probe <= e) lo += (u32)step;
    }
    const u32 ds = __shfl_sync(0xffffffffu, s, (int)lo);
    const u32 dend = __shfl_sync(0xffffffffu, inc, (int)lo);
    const u32 dsize = __shfl_sync(0xffffffffu, size, (int)lo);
    if (e < total) {
      const u32 j = e - (dend - dsize);
      bag_sufx[base + e] = (u32)SA[ds + j];
      bag_pos[base + e] = (ds + j) | (j == 0 ? BAG_HEAD : 0u);
    }
  }
}

__global__ void __launch_bounds__(256) k_bag_gather(const u32 *__restrict__ bag_sufx, u32 nb, const u32 *__restrict__ rank,
                                                    const u64 *__restrict__ state, const u32 *__restrict__ verdicts,
                                                    u32 round, u64 h, u32 n, u32 *__restrict__ r2out) {
  const u32 stride = gridDim.x * blockDim.x;
  const bool anyv = __ldg(verdicts) != 0u;
  for (u32 l = blockIdx.x * blockDim.x + threadIdx.x; l < nb; l += stride) {
    const u64 t = (u64)__ldg(bag_sufx + l) + h;
    u32 r2 = 0;
    if (t < n) {
      bool fin;
      const u32 w = __ldcg(rank + t);
      r2 = anyv ? resolve_label(w, state, round, &fin) : (w & RANK_MASK);
    }
    r2out[l] = r2;
  }
}

constexpr int BAG_THREADS = 512;
constexpr int BAG_TILE = BAG_THREADS - (int)TINY_MAX;

struct BagArgs {
  const u32 *sufx_in, *pos_in, *r2;
  u32 nb;
  u32 *sufx_out, *pos_out, *count_out;
  u32 *rank;
  i32 *SA;
};

__global__ void __launch_bounds__(BAG_THREADS) k_bag_refine(const BagArgs a) {
  constexpr u32 NW = BAG_THREADS / 32;
  __shared__ u32 s_sfx[BAG_THREADS], s_r2[BAG_THREADS];
  __shared__ u32 s_slot[BAG_THREADS];  // new slot, bit 31 = survives (its run has more than one member)
  __shared__ u32 s_gs[BAG_THREADS];    // slot of the entry (first slot of the group at a head)
  __shared__ u32 s_headm[NW];          // per warp: which entries start a group
  __shared__ u32 s_warp[NW], s_bal[NW];  // survivors per warp: count, lane mask
  __shared__ u32 s_base;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 l = blockIdx.x * (u32)BAG_TILE + tid;
  const bool have = l < a.nb;
  const u32 pos = have ? a.pos_in[l] : BAG_HEAD;  // past the end: looks like the start of another group
  s_sfx[tid] = have ? a.sufx_in[l] : 0u;
  s_r2[tid] = have ? a.r2[l] : 0u;
  s_gs[tid] = pos & ~BAG_HEAD;
  const u32 hm = __ballot_sync(0xffffffffu, (pos & BAG_HEAD) != 0u);
  if (lane == 0) s_headm[warp] = hm;
  __syncthreads();
  // my group: [g0, g1) in block coordinates; it belongs to this block iff its head is below BAG_TILE
  u32 g0 = 0, g1 = (u32)BAG_THREADS;
  bool head_seen;
  {
    u32 m = s_headm[warp] & (0xffffffffu >> (31u - lane));  // heads at or before me
    u32 w = warp;
    while (m == 0u && w > 0u) m = s_headm[--w];
    head_seen = m != 0u;  // false: the head is in the previous block
    if (head_seen) g0 = w * 32u + 31u - (u32)__clz(m);
    m = (lane == 31u) ? 0u : (s_headm[warp] & (0xffffffffu << (lane + 1u)));  // heads after me
    w = warp;
    while (m == 0u && w + 1u < NW) m = s_headm[++w];
    if (m != 0u) g1 = w * 32u + (u32)__ffs(m) - 1u;
  }
  const bool mine = have && head_seen && g0 < (u32)BAG_TILE;
  u32 less = 0, eq = 0, eq_before = 0;
  const u32 my = s_r2[tid];
  if (mine) {
    for (u32 m = g0; m < g1; ++m) {
      const u32 v = s_r2[m];
      less += (v < my) ? 1u : 0u;
      eq += (v == my) ? 1u : 0u;
      eq_before += (v == my && m < tid) ? 1u : 0u;
    }
  }
  const u32 gs = s_gs[g0];                        // first slot of the old group
  const u32 s1 = gs + less, slot = s1 + eq_before;  // first slot of my run, my own slot
  const bool surv = mine && eq > 1u;
  s_slot[tid] = mine ? (slot | (surv ? BAG_HEAD : 0u)) : 0u;
  // output position: survivors of earlier groups of the block (block order), then the survivors of
  // my own group in slot order -- the new groups stay contiguous, each headed by its first slot
  const u32 bal = __ballot_sync(0xffffffffu, surv);
  if (lane == 0u) { s_warp[warp] = (u32)__popc(bal); s_bal[warp] = bal; }
  __syncthreads();
  // The reservation of the block's output range is one atomic on a counter shared by the whole grid;
  // everything that does not need its result is done while it is in flight.
  if (tid == 0u) {
    u32 tot = 0;
#pragma unroll
    for (u32 w = 0; w < NW; ++w) tot += s_warp[w];
    s_base = tot ? atomicAdd(a.count_out, tot) : 0u;
  }
  const u32 sufx = s_sfx[tid];
  u32 o = 0;
  if (mine && !surv) {
    a.SA[slot] = (i32)sufx;
    a.rank[sufx] = RANK_DEAD | (slot + 1u);
  } else if (surv) {
    u32 in_group_before = 0;
    for (u32 m = g0; m < g1; ++m) {
      const u32 sm = s_slot[m];
      in_group_before += ((sm & BAG_HEAD) && (sm & ~BAG_HEAD) < slot) ? 1u : 0u;
    }
    if (tiny_label(s1) != tiny_label(gs)) a.rank[sufx] = tiny_label(s1);
    u32 before_group = (u32)__popc(s_bal[g0 >> 5] & ((1u << (g0 & 31u)) - 1u));  // survivors in front of my group's head
    for (u32 w = 0; w < (g0 >> 5); ++w) before_group += s_warp[w];
    o = before_group + in_group_before;
  }
  __syncthreads();
  if (surv) {
    o += s_base;
    a.sufx_out[o] = sufx;
    a.pos_out[o] = slot | (slot == s1 ? BAG_HEAD : 0u);
  }
}

// One launch instead of a dozen memsets: zero the counters selected by `mask` (bit w = word w of the
// counter block) and the digit histograms.
__global__ void __launch_bounds__(256) k_round_reset(u32 *__restrict__ ctr, u64 mask, u32 *__restrict__ ghist) {
  const u32 t = threadIdx.x;
  if (t < 64 && ((mask >> t) & 1ull)) ctr[t] = 0;
  for (u32 i = t; i < MAX_PASSES * RADIX; i += 256) ghist[i] = 0;
}

// A round that sorts at most SMALL_SORT elements does so in one block (bitonic network in shared
// memory, ordered by (key, suffix)) instead of launching up to eight radix passes, each with its
// look-back status reset: the late rounds of most texts sort a handful of suffixes, and there
// the launches are the cost.
constexpr u32 SMALL_SORT = 2048;
__global__ void __launch_bounds__(1024) k_small_sort(u64 *__restrict__ keys, u32 *__restrict__ vals, u32 S) {
  __shared__ u64 sk[SMALL_SORT];
  __shared__ u32 sv[SMALL_SORT];
  const u32 t = threadIdx.x;
  for (u32 i = t; i < SMALL_SORT; i += 1024) {
    sk[i] = (i < S) ? keys[i] : ~0ull;
    sv[i] = (i < S) ? vals[i] : 0xffffffffu;  // padding sorts behind every real element
  }
  __syncthreads();
  for (u32 k = 2; k <= SMALL_SORT; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      for (u32 i = t; i < SMALL_SORT; i += 1024) {
        const u32 x = i ^ j;
        if (x > i) {
          const bool up = (i & k) == 0;
          const u64 ka = sk[i], kb = sk[x];
          const u32 va = sv[i], vb = sv[x];
          const bool gt = ka > kb || (ka == kb && va > vb);
          if (gt == up) { sk[i] = kb; sk[x] = ka; sv[i] = vb; sv[x] = va; }
        }
      }
      __syncthreads();
    }
  }
  for (u32 i = t; i < S; i += 1024) { keys[i] = sk[i]; vals[i] = sv[i]; }
}

// ------------------------------------------------------------------------------------
// Host driver
// ------------------------------------------------------------------------------------
namespace {

constexpr int PASS_THREADS = 256;
constexpr int PASS_IPT = 16;
constexpr int PASS_TILE = PASS_THREADS * PASS_IPT;
#ifndef GSA_PASS_MIN_BLOCKS
#define GSA_PASS_MIN_BLOCKS 3
#endif
constexpr int PASS_MIN_BLOCKS = GSA_PASS_MIN_BLOCKS;  // resident CTAs per SM the pass kernel is compiled for
constexpr int RB_THREADS = 512;
constexpr int RB_IPT = 8;
constexpr int RB_TILE = RB_THREADS * RB_IPT;
constexpr int HIST_THREADS = 512;
#ifndef GSA_GA_THREADS
#define GSA_GA_THREADS 512
#define GSA_GA_IPT 8
#define GSA_GA_BLOCKS 2
#endif
constexpr int GA_THREADS = GSA_GA_THREADS;
constexpr int GA_IPT = GSA_GA_IPT;
constexpr int GA_BLOCKS_PER_SM = GSA_GA_BLOCKS;

struct Layout {
  u64 *packed; u64 packed_words;
  u64 *keys[2]; u32 *vals[2]; u32 *slots; u32 *lst[2]; u32 *rank;
  u64 *G;                      // [n + 2] slot range of every live group, indexed by label
  u64 *state, *rho, *rep;      // [n / HUGE_M + 2] per huge label (rep: HUGE_REPS entries each)
  u32 *seen;                   // [n / HUGE_M + 2] round in which k_huge_prepare last saw the label
  u64 *hkt_keys; u32 *hkt_labels; u32 hkt_cap;  // round-0 key -> label of the huge groups
  u32 *hfull;                  // [256] window histogram of round 0
  u32 *hlist[2]; u32 hcap;     // labels of the huge groups
  u32 *ctr;                    // [64] every counter the host reads back, in one block (one copy per sync point)
  u32 *hcount;                 // ctr + 4: [0], [1] entries of the two huge lists, [2] verdict flag of the round
  u32 *gupd; u32 *gupd_count;  // end-of-inert-block updates of one round (count: ctr + 16)
  u32 *tile_rtail, *next_rtail;
  u32 *ghist;      // [MAX_PASSES][256]
  u32 *bin_base;   // [MAX_PASSES][256]
  u32 *present;    // [256]
  u32 *skip_mask;  // [1]
  u32 *live_counter;  // [1] k_gather
  u32 *survivors;     // [1] k_tail_summary
  u32 *pass_status; size_t pass_status_words;  // counter at word 0 (256-word header), then [tiles][256]
  u32 *tile_tail, *next_tail, *tile_head, *prev_head;  // per rebuild tile
  u32 *bag_sufx[2], *bag_pos[2];               // the bag: (suffix, slot | BAG_HEAD), group after group
  u32 *bag_count;                              // [0], [1] entries of the two bag buffers, [2] descriptors
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
  y.slots = c.take<u32>(N);
  y.G = c.take<u64>(N + 2);
  y.state = c.take<u64>(N / HUGE_M + 2);
  y.rho = c.take<u64>(N / HUGE_M + 2);
  y.rep = c.take<u64>((N / HUGE_M + 2) * HUGE_REPS);
  y.seen = c.take<u32>(N / HUGE_M + 2);
  y.hkt_cap = 1024;
  while (y.hkt_cap < 4 * (N / HUGE_T + 1)) y.hkt_cap <<= 1;
  y.hkt_keys = c.take<u64>(y.hkt_cap);
  y.hkt_labels = c.take<u32>(y.hkt_cap);
  y.hfull = c.take<u32>(RADIX);
  y.hcap = (u32)(2 * (N / HUGE_T) + 4096);
  y.hlist[0] = c.take<u32>(y.hcap); y.hlist[1] = c.take<u32>(y.hcap);
  y.ctr = c.take<u32>(64);
  y.hcount = y.ctr + 4;
  y.gupd = c.take<u32>(3 * (size_t)y.hcap);
  y.gupd_count = y.ctr + 16;
  y.lst[0] = c.take<u32>(N);  y.lst[1] = c.take<u32>(N);
  y.rank = c.take<u32>(N);
  y.ghist = c.take<u32>(MAX_PASSES * RADIX);
  y.bin_base = c.take<u32>(MAX_PASSES * RADIX);
  y.present = c.take<u32>(256);
  y.skip_mask = y.ctr + 10;
  y.live_counter = y.ctr + 8;   // 64-bit word (live, to sort); ctr + 12: candidates of the pre-filter
  y.survivors = y.ctr + 14;
  const size_t ptiles = div_up(N, 2048);  // smallest tile of the pass configurations
  y.pass_status_words = 256 + ptiles * RADIX;
  y.pass_status = c.take<u32>(y.pass_status_words);
  const size_t rtiles = div_up(N, RB_TILE);
  y.tile_tail = c.take<u32>(rtiles + 1);
  y.next_tail = c.take<u32>(rtiles + 1);
  y.tile_head = c.take<u32>(rtiles + 1);
  y.prev_head = c.take<u32>(rtiles + 1);
  y.tile_rtail = c.take<u32>(rtiles + 1);
  y.next_rtail = c.take<u32>(rtiles + 1);
  for (int i = 0; i < 2; ++i) { y.bag_sufx[i] = c.take<u32>(N); y.bag_pos[i] = c.take<u32>(N); }
  y.bag_count = y.ctr + 0;
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

// Launch of the persistent pass: as many CTAs as fit on the device at once (never more than tiles).
template <int THREADS, int IPT, bool GEN, int MIN_BLOCKS>
static void launch_pass_p(const PassArgs &a, u32 L, cudaStream_t st) {
  static int per_sm = 0, sms = 0;  // (benign race: every thread computes the same values)
  const int smem = (int)PassPCfg<THREADS, IPT>::SMEM;
  if (per_sm == 0) {
    cudaFuncSetAttribute(k_radix_pass_p<THREADS, IPT, GEN, MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int dev = 0, n = 0, b = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_radix_pass_p<THREADS, IPT, GEN, MIN_BLOCKS>, THREADS, smem);
    sms = n > 0 ? n : kDefaultSMs;
    per_sm = b > 0 ? b : 1;
  }
  const u32 tiles = (u32)div_up(L, THREADS * IPT);
  const u32 grid = std::min<u32>(tiles, (u32)(sms * per_sm));
  k_radix_pass_p<THREADS, IPT, GEN, MIN_BLOCKS><<<grid, THREADS, smem, st>>>(a, tiles);
}

// Runs the radix passes for digits [0, npass) on `L` elements whose histograms are already
// in y.ghist.  `cur` is the buffer index holding the input (ignored when gen != null: the
// first pass then generates the keys and writes buffer 0).  *cur_out = buffer holding the
// sorted pairs.  k_scan_hist (bin offsets + constant digits) has been launched by the caller.
// `final_vals` (round 0): the LAST pass writes its values (the suffixes in sorted order) there instead of the ping-pong buffer --
// the caller passes the SA itself, which round 0 would otherwise fill with a copy of exactly that sequence.
static int run_passes(const Layout &y, u32 L, int npass, int cur, const KeyGen *gen, cudaStream_t st,
                      gsa_build_stats *stats, PassTimer &timer, int *cur_out, u32 *passes_done, u32 skip, int sms,
                      u32 *final_vals = nullptr, const u32 **vals_sorted = nullptr) {
  // `skip`: bit p = digit p is the same in every key (k_scan_hist), its pass would be the identity
  const u32 tiles = (u32)div_up(L, PASS_TILE);
  const size_t smem = PassCfg<PASS_THREADS, PASS_IPT>::SMEM;
  const char *cfg_env = getenv("GSA_PASS_CFG");  // experiments: alternative tile shapes of the pass kernel
  const int pass_cfg = cfg_env ? atoi(cfg_env) : 0;
  const u32 tiles_max = (u32)div_up(L, 2048);
  // L2 prefetch distance of the pass kernel in tiles: one generation of resident CTAs ahead (a CTA asks L2 for the tile
  // that its SM slot will most likely process next).  rep_1G: pass 0.576 -> 0.611 of the copy bandwidth, 268.8 -> 264.3 ms;
  // half that distance does the same, twice that distance nothing (profiles/r2/README.md).  GSA_PASS_PF=0 switches it off.
  const char *pf_env = getenv("GSA_PASS_PF");
  const u32 pf_dist = pf_env ? (u32)atoi(pf_env) : (u32)(sms * PASS_MIN_BLOCKS);
  bool need_gen = gen != nullptr;
  u32 done = 0;
  int last_p = -1;  // the last pass that will run
  {
    bool ng = need_gen;
    for (int p = 0; p < npass; ++p) {
      if (((skip >> p) & 1u) && !(ng && p == npass - 1)) continue;
      last_p = p;
      ng = false;
    }
  }
  const u32 *vals_now = gen ? nullptr : y.vals[cur];
  for (int p = 0; p < npass; ++p) {
    if (((skip >> p) & 1u) && !(need_gen && p == npass - 1)) continue;  // constant digit: identity pass
    u32 *const vdst_alt = (final_vals != nullptr && p == last_p) ? final_vals : nullptr;
    GSA_TRY(cudaMemsetAsync(y.pass_status, 0, (256 + (size_t)(pass_cfg ? tiles_max : tiles) * RADIX) * sizeof(u32), st));
    PassArgs a;
    a.n = L;
    a.shift = 8u * (u32)p;
    a.bin_base = y.bin_base + p * RADIX;
    a.counter = y.pass_status;
    a.status = y.pass_status + 256;
    a.pf_dist = pf_dist;
    cudaEvent_t t0, t1;
    GSA_TRY_RC(timer.next(&t0));
    GSA_TRY_RC(timer.next(&t1));
    GSA_TRY(cudaEventRecord(t0, st));
    if (need_gen) {
      a.keys_in = nullptr; a.vals_in = nullptr;
      a.keys_out = y.keys[0]; a.vals_out = vdst_alt ? vdst_alt : y.vals[0];
      a.gen = *gen;
      if (pass_cfg == 10) launch_pass_p<512, 8, true, 2>(a, L, st);
      else if (pass_cfg == 11) launch_pass_p<256, 16, true, 2>(a, L, st);
      else k_radix_pass<PASS_THREADS, PASS_IPT, true><<<tiles, PASS_THREADS, smem, st>>>(a);
      cur = 0;
      need_gen = false;
    } else {
      a.keys_in = y.keys[cur]; a.vals_in = y.vals[cur];
      a.keys_out = y.keys[cur ^ 1]; a.vals_out = vdst_alt ? vdst_alt : y.vals[cur ^ 1];
      a.gen = KeyGen{};
      if (pass_cfg == 1) {
        k_radix_pass<256, 12, false, 4><<<(u32)div_up(L, 256 * 12), 256, PassCfg<256, 12>::SMEM, st>>>(a);
      } else if (pass_cfg == 2) {
        k_radix_pass<384, 16, false, 2><<<(u32)div_up(L, 384 * 16), 384, PassCfg<384, 16>::SMEM, st>>>(a);
      } else if (pass_cfg == 3) {
        k_radix_pass<512, 12, false, 2><<<(u32)div_up(L, 512 * 12), 512, PassCfg<512, 12>::SMEM, st>>>(a);
      } else if (pass_cfg == 10) {
        launch_pass_p<512, 8, false, 2>(a, L, st);
      } else if (pass_cfg == 11) {
        launch_pass_p<256, 16, false, 2>(a, L, st);
      } else if (pass_cfg == 12) {
        launch_pass_p<384, 6, false, 3>(a, L, st);
      } else if (pass_cfg == 13) {
        launch_pass_p<256, 8, false, 4>(a, L, st);
      } else {
        k_radix_pass<PASS_THREADS, PASS_IPT, false, PASS_MIN_BLOCKS><<<tiles, PASS_THREADS, smem, st>>>(a);
      }
      cur ^= 1;
    }
    KLAUNCH_CHECK();
    vals_now = a.vals_out;
    GSA_TRY(cudaEventRecord(t1, st));
    ++done;
    if (stats) {
      stats->radix_pass_launches++;
      stats->radix_pass_elements += L;
      // 12 B read + 12 B written per element; the key-generating first pass of round 0 reads b / 8 B of packed text instead
      stats->radix_pass_bytes += (gen != nullptr && done == 1) ? (u64)L * gen->b / 8 + (u64)L * 12 : (u64)L * 24;
      stats->kernel_launches++;
    }
  }
  if (stats) stats->kernel_launches++;  // k_scan_hist
  *cur_out = cur;
  if (vals_sorted) *vals_sorted = vals_now;
  *passes_done = done;
  return GSA_OK;
}

// 64 counter words + 256 alphabet words of page-locked host memory per calling thread (allocated once, portable across
// devices, released when the thread ends).
static u32 *host_mailbox() {
  struct Box {
    u32 *p = nullptr;
    ~Box() { if (p) cudaFreeHost(p); }
  };
  static thread_local Box box;
  if (box.p == nullptr && cudaHostAlloc(reinterpret_cast<void **>(&box.p), (64 + 256) * sizeof(u32), cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    box.p = nullptr;
  }
  return box.p;
}

int build_sa_device(const u8 *d_T, i32 *d_SA, u32 n, void *workspace, size_t workspace_bytes, cudaStream_t st,
                    gsa_build_stats *stats) {
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n == 0) return GSA_OK;
  static_assert(PASS_THREADS >= RADIX, "");
  int sms = kDefaultSMs;
  {
    // opt in to > 48 KB dynamic shared memory: once per device and process (seven driver calls otherwise paid by every build)
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    GSA_TRY(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit == 0ull || !(configured.load(std::memory_order_acquire) & bit)) {
    const int smem = (int)PassCfg<PASS_THREADS, PASS_IPT>::SMEM;
    GSA_TRY(cudaFuncSetAttribute(k_prefilter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM));
    GSA_TRY(cudaFuncSetAttribute(k_prefilter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM));
    GSA_TRY(cudaFuncSetAttribute(k_radix_pass<PASS_THREADS, PASS_IPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    GSA_TRY(cudaFuncSetAttribute(k_radix_pass<PASS_THREADS, PASS_IPT, false, PASS_MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    GSA_TRY(cudaFuncSetAttribute(k_radix_pass<256, 12, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PassCfg<256, 12>::SMEM));
    GSA_TRY(cudaFuncSetAttribute(k_radix_pass<384, 16, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PassCfg<384, 16>::SMEM));
    GSA_TRY(cudaFuncSetAttribute(k_radix_pass<512, 12, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PassCfg<512, 12>::SMEM));
    configured.fetch_or(bit, std::memory_order_release);
    }
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
  // Everything the host reads back lands in a small page-locked block owned by the calling thread: a copy into pageable
  // memory (a stack array) goes through the driver's staging path, ~10 us more per round trip, and a 4 MiB text pays six.
  u32 *const hostbox = host_mailbox();
  if (hostbox == nullptr) { set_error("cudaHostAlloc of the 2 KiB read-back block failed", __FILE__, __LINE__); return GSA_ENOMEM; }
  u32 *const present = hostbox + 64;
  GSA_TRY(cudaMemcpyAsync(present, y.present, 256 * sizeof(u32), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaStreamSynchronize(st));
  CodeMap cm;
  u32 sigma = 0;
  for (int c = 0; c < 256; ++c) {
    cm.code[c] = (u8)sigma;
    if (present[c]) ++sigma;
  }
  const u32 b = bits_for(sigma > 1 ? sigma - 1 : 1);  // codes 0..sigma-1
  // Symbols per round-0 key.  64 bits hold 64/b symbols, but sorting more than about
  // log2(n) + 10 bits only orders suffixes that are (for text without long repeats) already
  // unique: every 8 bits beyond that is a full radix pass over all n suffixes, while the few
  // groups that are still tied are cheaper to finish in a doubling round over just them.
  const u32 k_max = 64 / b;
  const u32 k_want = std::min<u32>(k_max, (bits_for(n) + 10 + b - 1) / b);
  const u64 nwords = ((u64)n * b + 63) / 64 + 2;
  {
    k_pack<<<(u32)div_up(nwords, 256), 256, 0, st>>>(d_T, n, b, cm, y.packed, nwords);
    KLAUNCH_CHECK();
  }
  u32 k = k_want;
  if (const char *k_env = getenv("GSA_KEY_SYMBOLS")) {  // experiments: force the round-0 depth
    k = (u32)atoi(k_env);
  } else if (k_want < k_max && n >= (1u << 23)) {  // (below 8 MiB the probe's host round trip costs more than a wrong depth)
    // ... unless the text is repetitive: then nearly everything survives round 0 whatever its
    // depth, and a deeper start means fewer doubling rounds.
    constexpr u32 M = 1u << 16, TBL = 1u << 18;
    u64 *table = reinterpret_cast<u64 *>(y.slots);
    GSA_TRY(cudaMemsetAsync(table, 0, TBL * sizeof(u64), st));
    GSA_TRY(cudaMemsetAsync(y.survivors, 0, sizeof(u32), st));
    KeyGen probe{y.packed, n, 0, b, k_want * b};
    k_sample_dups<<<M / 256, 256, 0, st>>>(probe, M, table, TBL - 1, y.survivors);
    KLAUNCH_CHECK();
    GSA_TRY(cudaMemcpyAsync(hostbox, y.survivors, sizeof(u32), cudaMemcpyDeviceToHost, st));
    GSA_TRY(cudaStreamSynchronize(st));
    const u32 dups = hostbox[0];
    if (dups > M / 8) k = k_max;
    // ... and if the sampled suffixes fall into very few classes the text consists of a few huge
    // groups: those are cheap to refine (their inert majority is not sorted), so a 32-bit start
    // saves half of the round-0 passes for the price of one cheap extra round.
    if (M - dups < M / 16 && !getenv("GSA_NO_INERT")) k = std::max<u32>(1, std::min<u32>(k_max, 32 / b));
    if (stats) stats->kernel_launches++;
  }
  if (k < 1) k = 1;
  if (k > k_max) k = k_max;
  if (8 % b == 0 && !getenv("GSA_KEY_SYMBOLS")) k = std::min<u32>(k_max, (k + 8 / b - 1) / (8 / b) * (8 / b));  // whole digits
  const u32 key_bits = k * b;
  const u32 ns = (k - 1 < n) ? (k - 1) : n;           // short suffixes
  if (stats) {
    stats->sigma = sigma; stats->bits_per_symbol = b; stats->symbols_per_key = k;
    stats->kernel_launches += 2;
  }

  // ---- round 0 ------------------------------------------------------------------------------
  KeyGen gen{y.packed, n, ns, b, key_bits};
  const int npass0 = (int)div_up(key_bits, 8);
  const u32 hist_blocks = (u32)std::min<u64>((u64)sms * 4, std::max<u64>(1, div_up(n, HIST_THREADS)));
  GSA_TRY(cudaMemsetAsync(y.ghist, 0, MAX_PASSES * RADIX * sizeof(u32), st));
  if (8 % b == 0 && key_bits % 8 == 0 && n > 64) {
    GSA_TRY(cudaMemsetAsync(y.hfull, 0, RADIX * sizeof(u32), st));
    k_hist0_windows<HIST_THREADS><<<hist_blocks, HIST_THREADS, 0, st>>>(y.packed, n, b, y.hfull);
    KLAUNCH_CHECK();
    k_hist0_finish<<<1, RADIX, 0, st>>>(y.packed, n, b, key_bits, npass0, y.hfull, y.ghist);
    KLAUNCH_CHECK();
    if (stats) stats->kernel_launches += 2;
  } else {
    k_hist0<HIST_THREADS><<<hist_blocks, HIST_THREADS, 0, st>>>(gen, npass0, y.ghist);
    KLAUNCH_CHECK();
    if (stats) stats->kernel_launches++;
  }
  int cur = 0;
  u32 passes = 0;
  const u32 *sufx0 = nullptr;  // round 0: where the sorted suffixes are (the SA itself unless GSA_NO_SA_ALIAS)
  PassTimer timer;
  // Host <-> device traffic of the control flow: every counter the host needs lives in one 64-word block
  // (y.ctr), read back with ONE copy per sync point (`mailbox`); the counters of a round are zeroed by one
  // kernel.  Sync points per doubling round: after the gather (how much is live / to be sorted, which digits
  // are constant) and at the end (survivors, bag, huge groups); large rounds add one before the rebuild (is
  // this the last round?), small ones do not bother.
  u32 *const mailbox = hostbox;
  auto fetch_counters = [&]() -> int {
    GSA_TRY(cudaMemcpyAsync(mailbox, y.ctr, 64 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    GSA_TRY(cudaStreamSynchronize(st));
    return GSA_OK;
  };
  constexpr u32 SYNC_WORTH = 1u << 20;  // below this many elements an extra host round trip costs more than it can save
  GSA_TRY(cudaMemsetAsync(y.ctr, 0, 64 * sizeof(u32), st));
  GSA_TRY(cudaEventRecord(ev[1], st));
  {
    u32 skip0 = 0;
    k_scan_hist<<<npass0, RADIX, 0, st>>>(y.ghist, y.bin_base, n, y.skip_mask);
    KLAUNCH_CHECK();
    if (n >= 8 * SYNC_WORTH) {  // a constant digit (a^n ...) saves a pass over all n suffixes: worth a round trip from 8 Mi on
      GSA_TRY_RC(fetch_counters());
      skip0 = mailbox[10];
    }
    GSA_TRY_RC(run_passes(y, n, npass0, 0, &gen, st, stats, timer, &cur, &passes, skip0, sms,
                          getenv("GSA_NO_SA_ALIAS") ? nullptr : reinterpret_cast<u32 *>(d_SA), &sufx0));
  }
  GSA_TRY(cudaEventRecord(ev[2], st));

  // key = label(i) << 32 | label(i + h).  The label half starts on a digit boundary on purpose:
  // huge labels are multiples of 256, so in a round that sorts only members of huge groups the
  // digit holding their low byte is constant and its pass is skipped (as are the digits above
  // bits_for(n) in both halves).
  const u32 lab_bits = 32;
  bool sparse = false;
  int hcur = 0;  // hlist[hcur] / hcount[hcur]: huge groups entering the next round
  GSA_TRY(cudaMemsetAsync(y.state, 0, ((size_t)n / HUGE_M + 2) * sizeof(u64), st));
  GSA_TRY(cudaMemsetAsync(y.rho, 0, ((size_t)n / HUGE_M + 2) * sizeof(u64), st));
  GSA_TRY(cudaMemsetAsync(y.seen, 0, ((size_t)n / HUGE_M + 2) * sizeof(u32), st));
  GSA_TRY(cudaMemsetAsync(y.rep, 0xff, ((size_t)n / HUGE_M + 2) * HUGE_REPS * sizeof(u64), st));
  // tail summaries + survivor count, then the rebuild proper.  When the survivor count is fetched before
  // the rebuild is launched, the last round can skip its rank writes.
  // The bag is off in sparse mode (most suffixes then carry no label at all).
  u32 tiny_conf = getenv("GSA_NO_BAG") ? 0u : TINY_MAX;
  if (const char *e = getenv("GSA_TINY_MAX")) tiny_conf = std::min<u32>(TINY_MAX, (u32)atoi(e));
  int bcur = 0;  // bag buffer the rebuild of the current round appends to (= input of the next round)
  bool list_used = false;  // round 0 went through k_rebuild_list
  RebuildArgs r0_args{};
  u32 r0_tiles = 0;
  auto launch_rebuild = [&](bool round0, u32 rnd, u32 L, int kv, bool may_finish) -> int {
    const u32 tiles = (u32)div_up(L, RB_TILE);
    GSA_TRY(cudaMemsetAsync(y.survivors, 0, sizeof(u32), st));  // (also the probe's / the sparse mode's scratch counter)
    RebuildArgs r;
    r.keys = y.keys[kv]; r.sufx = (round0 && sufx0 != nullptr) ? sufx0 : y.vals[kv];
    r.sa_holds_sufx = round0 && r.sufx == reinterpret_cast<const u32 *>(d_SA);
    r.pos_in = round0 ? nullptr : y.slots;
    r.L = L;
    r.short_from = round0 ? (n - ns) : 0xffffffffu;
    r.lab_bits = lab_bits;
    r.rank = y.rank; r.SA = d_SA;
    r.G = y.G; r.hlist = y.hlist[hcur]; r.hcount = y.hcount + hcur; r.rep = y.rep; r.round = rnd;
    r.hkt = HugeKeyTable{y.hkt_keys, y.hkt_labels, y.hkt_cap - 1};
    r.survivors = y.survivors;
    r.tile_tail = y.tile_tail; r.next_tail = y.next_tail;
    r.tile_head = y.tile_head; r.prev_head = y.prev_head;
    r.tiny_max = sparse ? 0u : tiny_conf;
    r.bag_desc = y.keys[kv ^ 1];  // the other half of the sort's double buffer is free until the next walk
    r.bag_desc_count = y.bag_count + 2;
    r.surv_list = y.lst[0]; r.surv_count = y.ctr + 18;  // (round 0, sparse mode only; the word is zero from the start)
    const bool list_live = round0 && !getenv("GSA_NO_LIVE_LIST");
    r.live_idx = list_live ? y.slots : nullptr;       // (free in round 0)
    r.live_idx_ctl = y.ctr + 19;                      // words 19, 20, 21: zero from the start
    if (round0) k_tail_summary<RB_THREADS, RB_IPT, true><<<tiles, RB_THREADS, 0, st>>>(r);
    else k_tail_summary<RB_THREADS, RB_IPT, false><<<tiles, RB_THREADS, 0, st>>>(r);
    KLAUNCH_CHECK();
    k_tail_scan<<<1, 1024, 0, st>>>(y.tile_tail, y.next_tail, tiles, y.tile_head, y.prev_head);
    KLAUNCH_CHECK();
    bool fin = false;
    u32 surv_bound = L;  // survivors <= L
    if (round0 || L >= SYNC_WORTH) {
      GSA_TRY_RC(fetch_counters());
      const u32 surv = mailbox[14], bag_left = mailbox[bcur];
      surv_bound = surv;
      fin = may_finish && surv == 0 && bag_left == 0;  // nothing is live after this round
      if (round0) {
        // few survivors: do not scatter n ranks for the sake of a handful of look-ups
        sparse = surv != 0 && (u64)surv * 64 < n && !getenv("GSA_NO_SPARSE");
        if (sparse) GSA_TRY(cudaMemsetAsync(y.rank, 0, (size_t)n * sizeof(u32), st));
        r.tiny_max = sparse ? 0u : tiny_conf;
      }
    }
    if (round0) {
      // sparse, every non-unique element listed by k_tail_summary: handle just those (the caller checks ctl[2] at its next
      // read-back and falls back to the general kernel if a large group was met)
      list_used = sparse && !fin && r.sa_holds_sufx && r.live_idx != nullptr && mailbox[20] == 0u && mailbox[19] == surv_bound;
      r0_args = r;
      r0_tiles = tiles;
      if (fin && r.sa_holds_sufx) {
        // everything is unique and the SA already holds the order: nothing left to write
      } else if (fin) {
        k_rebuild<RB_THREADS, RB_IPT, true, RB_FINAL><<<tiles, RB_THREADS, 0, st>>>(r);
      } else if (list_used) {
        k_rebuild_list<<<(u32)div_up(surv_bound, 256), 256, 0, st>>>(r, surv_bound);
      } else if (sparse) {
        k_rebuild<RB_THREADS, RB_IPT, true, RB_SPARSE><<<tiles, RB_THREADS, 0, st>>>(r);
      } else {
        k_rebuild<RB_THREADS, RB_IPT, true, RB_NORMAL><<<tiles, RB_THREADS, 0, st>>>(r);
      }
    } else {
      if (fin) k_rebuild<RB_THREADS, RB_IPT, false, RB_FINAL><<<tiles, RB_THREADS, 0, st>>>(r);
      else k_rebuild<RB_THREADS, RB_IPT, false, RB_NORMAL><<<tiles, RB_THREADS, 0, st>>>(r);
    }
    KLAUNCH_CHECK();
    if (stats) stats->kernel_launches += 3;
    if (r.tiny_max && surv_bound > 1) {  // at most survivors / 2 new tiny groups
      k_bag_append<<<(u32)div_up(surv_bound / 2, 256), 256, 0, st>>>(r.bag_desc, r.bag_desc_count, d_SA, y.bag_sufx[bcur],
                                                                    y.bag_pos[bcur], y.bag_count + bcur);
      KLAUNCH_CHECK();
      if (stats) stats->kernel_launches++;
    }
    return GSA_OK;
  };

  GSA_TRY(cudaMemsetAsync(y.hkt_labels, 0, (size_t)y.hkt_cap * sizeof(u32), st));
  GSA_TRY_RC(launch_rebuild(true, 0, n, cur, true));
  GSA_TRY_RC(fetch_counters());  // end of round 0: survivors, bag entries, huge groups
  if (list_used && mailbox[21] != 0u) {  // a group too large for k_rebuild_list: the general kernel redoes round 0's rebuild
    GSA_TRY(cudaMemsetAsync(y.ctr + 18, 0, sizeof(u32), st));
    k_rebuild<RB_THREADS, RB_IPT, true, RB_SPARSE><<<r0_tiles, RB_THREADS, 0, st>>>(r0_args);
    KLAUNCH_CHECK();
    GSA_TRY_RC(fetch_counters());
  }
  u32 survivors = mailbox[14];
  u32 nbag = mailbox[bcur];     // entries of bag buffer bcur
  u32 hc = mailbox[4 + hcur];   // huge groups entering round 1
  if (hc) {  // label the members of the huge groups in text order
    const u32 blocks = (u32)std::min<u64>((u64)sms * 8, std::max<u64>(1, div_up(n, 256)));
    k_rank_huge0<<<blocks, 256, 0, st>>>(gen, HugeKeyTable{y.hkt_keys, y.hkt_labels, y.hkt_cap - 1}, n - ns, y.rank);
    KLAUNCH_CHECK();
    if (stats) stats->kernel_launches++;
  }
  GSA_TRY(cudaEventRecord(ev[3], st));
  u32 round = 0;
  auto log_round = [&](u64 depth, u64 live, u32 sorted, u32 groups, u32 kb, u32 np, u32 bag) {
    if (!stats) return;
    if (round >= GSA_MAX_ROUNDS) return;
    gsa_round_stat &s = stats->round[round];
    s.depth = depth; s.live = live; s.sorted = sorted; s.groups = groups; s.key_bits = kb; s.passes = np; s.bag = bag;
    cudaEventElapsedTime(&s.ms_total, ev[0], ev[3]);
    cudaEventElapsedTime(&s.ms_sort, ev[1], ev[2]);
    stats->rounds = round + 1;
  };
  GSA_TRY(cudaEventSynchronize(ev[3]));
  if (stats) stats->ms_radix_passes += timer.drain();
  log_round(k, n, n, 0, key_bits, passes, 0);

  // ---- doubling rounds ------------------------------------------------------------------------
  int lcur = 0;        // lst buffer holding the candidates (text order); round 1 uses 0..n-1 ...
  u32 Lcand = n;       // candidates to walk
  u32 live = survivors;
  bool ident = true;
  if (sparse && !getenv("GSA_NO_SURV_LIST")) {  // ... unless round 0 listed its few survivors (k_rebuild<RB_SPARSE>, any order)
    Lcand = survivors;
    ident = false;
  }
  u64 h = k;           // suffixes are sorted by their first h symbols
  const u32 kb = 2 * lab_bits;
  const int npass = (int)div_up(kb, 8);
  const bool filter_allowed = !getenv("GSA_NO_INERT");
  const bool use_prefilter = !getenv("GSA_NO_PREFILTER");
  const bool use_small_sort = !getenv("GSA_NO_SMALL_SORT");
  const u32 tiny_max = sparse ? 0u : tiny_conf;
  while (live > 0 || nbag > 0) {
    ++round;
    GSA_TRY(cudaEventRecord(ev[0], st));
    // at most n / HUGE_T groups survive k_huge_prepare (duplicates dropped) and at most as many are appended by a
    // rebuild, so the list cannot outgrow its 2 n / HUGE_T + 4096 entries; anything else is a bug, not an input
    if (hc >= y.hcap) { set_error("huge-group list overflow", __FILE__, __LINE__); return GSA_ECUDA; }
    const int bin = bcur;
    bcur ^= 1;
    // counters of this round: the new huge list, verdict flag, live / sort counts, skip mask, pre-filter candidates,
    // survivors, inert-block updates, bag descriptors, the bag buffer this round fills; + the digit histograms
    {
      const u64 mask = (1ull << (4 + (hcur ^ 1))) | (1ull << 6) | (1ull << 8) | (1ull << 9) | (1ull << 10) | (1ull << 12) |
                       (1ull << 14) | (1ull << 16) | (1ull << 2) | (1ull << bcur);
      k_round_reset<<<1, 256, 0, st>>>(y.ctr, mask, y.ghist);
      KLAUNCH_CHECK();
    }
    // huge groups: verdicts for this round, and the list for the next one
    if (hc) {
      k_huge_prepare<<<(u32)div_up(hc, 256), 256, 0, st>>>(y.hlist[hcur], hc, y.G, y.state, y.rep, round, tiny_max, y.hlist[hcur ^ 1], y.hcount + (hcur ^ 1), y.hcount + 2, getenv("GSA_NO_DEDUPE") ? nullptr : y.seen);
      KLAUNCH_CHECK();
      if (stats) stats->kernel_launches++;
    }
    hcur ^= 1;  // the rebuild of this round appends the huge groups it creates to the new list
    const bool filter = filter_allowed && !sparse && hc > 0;
    if (filter) {
      k_huge_rho<<<(u32)div_up(hc, 256), 256, 0, st>>>(y.hlist[hcur], y.hcount + hcur, y.rank, y.state, y.rep, y.rho, round, h, n);
      KLAUNCH_CHECK();
      if (stats) stats->kernel_launches++;
    }
    GatherArgs g;
    g.lst_in = ident ? nullptr : y.lst[lcur];
    g.Lin = Lcand; g.rank = y.rank; g.SA = d_SA; g.n = n; g.h = h; g.lab_bits = lab_bits;
    g.npass = sparse ? 0 : npass;  // sparse: keys are completed by k_lazy_fill, histogram afterwards
    g.tiny_max = tiny_max;
    g.round = round; g.filter = filter ? 1 : 0; g.state = y.state; g.verdicts = y.hcount + 2; g.rho = y.rho; g.rep = y.rep;
    // While most of the text is live the candidate list is not worth its 8 bytes per suffix:
    // the next round simply walks all text positions again.
    const bool write_list = (u64)live * 2 < n || sparse;
    g.keys_out = y.keys[0]; g.vals_out = y.vals[0]; g.lst_out = write_list ? y.lst[lcur ^ 1] : nullptr;
    g.counter = y.live_counter; g.ghist = y.ghist;
    g.sa0 = sparse ? d_SA : nullptr;
    g.gen = gen;
    g.todo = y.slots;  // free until k_slots of this round
    g.todo_count = y.survivors;
    g.Lin_ptr = nullptr;
    // dense round: settle the inert majority in a streaming pre-filter, k_gather sees only the rest
    if (ident && !write_list && filter && !sparse && use_prefilter && hc <= PF_MAX_GROUPS) {
      u32 *cand_count = y.live_counter + 4;
      PrefilterArgs pf;
      pf.rank = y.rank; pf.n = n; pf.h = h; pf.round = round; pf.tiny_max = tiny_max;
      pf.state = y.state; pf.verdicts = y.hcount + 2; pf.rho = y.rho; pf.rep = y.rep;
      pf.hlist = y.hlist[hcur]; pf.hcount = y.hcount + hcur;
      pf.cand = y.lst[0]; pf.cand_count = cand_count; pf.counter = y.live_counter;
      const u32 pblocks = (u32)std::min<u64>((u64)sms * PF_BLOCKS_PER_SM, std::max<u64>(1, div_up(n, PF_WCHUNK * PF_WARPS)));
      if (h % 4 == 0) k_prefilter<true><<<pblocks, PF_THREADS, PF_SMEM, st>>>(pf);
      else k_prefilter<false><<<pblocks, PF_THREADS, PF_SMEM, st>>>(pf);
      KLAUNCH_CHECK();
      if (stats) stats->kernel_launches++;
      g.lst_in = y.lst[0];
      g.Lin_ptr = cand_count;
    }
    const u32 gblocks = (u32)std::min<u64>((u64)sms * GA_BLOCKS_PER_SM, std::max<u64>(1, div_up(Lcand, GA_THREADS * GA_IPT)));
    k_gather<GA_THREADS, GA_IPT, GA_BLOCKS_PER_SM><<<gblocks, GA_THREADS, 0, st>>>(g);
    KLAUNCH_CHECK();
    if (stats) stats->kernel_launches += 2;
    // the bag: refine the tiny groups.  Every label read of the round (k_gather, k_bag_gather) precedes
    // every label write (k_bag_refine, k_rebuild), so all readers see the labels left by the last round.
    if (nbag) {
      const u32 bb = (u32)std::min<u64>((u64)sms * 8, div_up(nbag, 256));
      k_bag_gather<<<bb, 256, 0, st>>>(y.bag_sufx[bin], nbag, y.rank, y.state, y.hcount + 2, round, h, n, y.slots);
      KLAUNCH_CHECK();
      BagArgs ba;
      ba.sufx_in = y.bag_sufx[bin]; ba.pos_in = y.bag_pos[bin]; ba.r2 = y.slots; ba.nb = nbag;
      ba.sufx_out = y.bag_sufx[bcur]; ba.pos_out = y.bag_pos[bcur]; ba.count_out = y.bag_count + bcur;
      ba.rank = y.rank; ba.SA = d_SA;
      k_bag_refine<<<(u32)div_up(nbag, BAG_TILE), BAG_THREADS, 0, st>>>(ba);
      KLAUNCH_CHECK();
      if (stats) stats->kernel_launches += 2;
    }
    // bin offsets and constant digits from the histogram the gather left behind; the element count is still on the device
    if (!sparse) {
      k_scan_hist<<<npass, RADIX, 0, st>>>(y.ghist, y.bin_base, 0, y.skip_mask, y.live_counter + 1);
      KLAUNCH_CHECK();
    }
    GSA_TRY_RC(fetch_counters());
    const u32 Llive = mailbox[8], S = mailbox[9];
    u32 skip = mailbox[10];
    passes = 0;
    GSA_TRY(cudaEventRecord(ev[1], st));
    GSA_TRY(cudaEventRecord(ev[2], st));
    if (S > 0) {
      if (sparse) {
        // at most one entry per sorted suffix (S * 64 < n, so the pairs fit in the slot buffer)
        k_lazy_fill<<<(u32)div_up(S, 256), 256, 0, st>>>(gen, d_SA, g.todo, g.todo_count, y.keys[0]);
        KLAUNCH_CHECK();
        if (stats) stats->kernel_launches += 1;
        skip = 0;
        if (!use_small_sort || S > SMALL_SORT) {
          const u32 hb = (u32)std::min<u64>((u64)sms * 4, std::max<u64>(1, div_up(S, HIST_THREADS)));
          k_hist_keys<HIST_THREADS><<<hb, HIST_THREADS, 0, st>>>(y.keys[0], S, npass, y.ghist);
          KLAUNCH_CHECK();
          k_scan_hist<<<npass, RADIX, 0, st>>>(y.ghist, y.bin_base, S, y.skip_mask);
          KLAUNCH_CHECK();
          if (stats) stats->kernel_launches += 2;
          if (S >= SYNC_WORTH) {
            GSA_TRY_RC(fetch_counters());
            skip = mailbox[10];
          }
        }
      }
      GSA_TRY(cudaEventRecord(ev[1], st));
      if (use_small_sort && S <= SMALL_SORT) {
        k_small_sort<<<1, 1024, 0, st>>>(y.keys[0], y.vals[0], S);
        KLAUNCH_CHECK();
        cur = 0;
        if (stats) stats->kernel_launches++;
      } else {
        GSA_TRY_RC(run_passes(y, S, npass, 0, nullptr, st, stats, timer, &cur, &passes, skip, sms));
      }
      GSA_TRY(cudaEventRecord(ev[2], st));
      // SA slots of the sorted elements from the group tables
      const u32 tiles = (u32)div_up(S, RB_TILE);
      SlotArgs sa;
      sa.keys = y.keys[cur]; sa.S = S; sa.lab_bits = lab_bits; sa.G = y.G; sa.rho = y.rho; sa.round = round;
      sa.filter = filter ? 1 : 0; sa.slots = y.slots;
      sa.tile_rtail = y.tile_rtail; sa.next_rtail = y.next_rtail; sa.tile_rhead = y.tile_head; sa.prev_rhead = y.prev_head; sa.gupd = y.gupd; sa.gupd_count = y.gupd_count;
      k_run_summary<RB_THREADS, RB_IPT><<<tiles, RB_THREADS, 0, st>>>(sa);
      KLAUNCH_CHECK();
      k_tail_scan<<<1, 1024, 0, st>>>(y.tile_rtail, y.next_rtail, tiles, y.tile_head, y.prev_head);
      KLAUNCH_CHECK();
      k_slots<RB_THREADS, RB_IPT><<<tiles, RB_THREADS, 0, st>>>(sa);
      KLAUNCH_CHECK();
      if (filter) {
        k_apply_g<<<(u32)div_up(2 * (u64)hc + 2, 256), 256, 0, st>>>(y.gupd, y.gupd_count, y.G);
        KLAUNCH_CHECK();
      }
      if (stats) stats->kernel_launches += 4;
      GSA_TRY_RC(launch_rebuild(false, round, S, cur, Llive == S));
    }
    GSA_TRY(cudaEventRecord(ev[3], st));
    GSA_TRY_RC(fetch_counters());  // end of the round: survivors, the bag, the huge groups of the next round
    survivors = (S > 0) ? mailbox[14] : 0u;
    const u32 hc_this = hc;
    hc = mailbox[4 + hcur];
    h *= 2;
    if (stats) stats->ms_radix_passes += timer.drain();
    log_round(h, Llive, S, hc_this, kb, passes, nbag);
    live = (Llive - S) + survivors;  // inert members + non-unique sorted ones (some of them now in the bag)
    nbag = mailbox[bcur];
    if (write_list) {
      lcur ^= 1;
      Lcand = Llive;
      ident = false;
    }
    if (round > 64) { set_error("prefix doubling did not converge", __FILE__, __LINE__); return GSA_ECUDA; }
  }
  GSA_TRY(cudaEventRecord(ev_all1, st));
  GSA_TRY(cudaStreamSynchronize(st));
  if (stats) cudaEventElapsedTime(&stats->ms_total, ev_all0, ev_all1);
  return GSA_OK;
}

// ------------------------------------------------------------------------------------
// Stable order of a byte string: order[q] = index of the q-th byte in (value, index) order.
// One onesweep pass of the radix kernel above, its keys generated from the packed bytes; used by
// the inverse BWT (bwt.cu), where this order is the LF/psi permutation of the transform.
// ------------------------------------------------------------------------------------
namespace {
struct ByteOrderLayout { u64 *packed; u64 nwords; u64 *keys_out; u32 *ghist, *bin_base, *skip_mask, *pass_status; size_t total; };
ByteOrderLayout byte_order_layout(char *base, u32 n) {
  ByteOrderLayout y;
  Carve c{base, 0};
  y.nwords = (u64)n / 8 + 4;
  y.packed = c.take<u64>(y.nwords);
  y.keys_out = c.take<u64>(n);
  y.ghist = c.take<u32>(MAX_PASSES * RADIX);
  y.bin_base = c.take<u32>(MAX_PASSES * RADIX);
  y.skip_mask = c.take<u32>(64);
  y.pass_status = c.take<u32>(256 + div_up(n, PASS_TILE) * RADIX);
  y.total = c.used;
  return y;
}
}  // namespace

size_t byte_order_workspace_bytes(u32 n) { return byte_order_layout(nullptr, n == 0 ? 1 : n).total + 256; }

int stable_byte_order_device(const u8 *d_bytes, u32 n, u32 *d_order, u32 *d_counts, void *workspace, size_t workspace_bytes,
                             cudaStream_t st) {
  if (n == 0) return GSA_OK;
  if (workspace == nullptr || workspace_bytes < byte_order_workspace_bytes(n)) {
    set_error("workspace too small", __FILE__, __LINE__);
    return GSA_EINVAL;
  }
  const size_t mis = (256 - (reinterpret_cast<uintptr_t>(workspace) & 255)) & 255;
  const ByteOrderLayout y = byte_order_layout(static_cast<char *>(workspace) + mis, n);
  const int smem = (int)PassCfg<PASS_THREADS, PASS_IPT>::SMEM;
  GSA_TRY(cudaFuncSetAttribute(k_radix_pass<PASS_THREADS, PASS_IPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int sms = kDefaultSMs, dev = 0;
  GSA_TRY(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CodeMap cm;
  for (int c = 0; c < 256; ++c) cm.code[c] = (u8)c;
  k_pack<<<(u32)div_up(y.nwords, 256), 256, 0, st>>>(d_bytes, n, 8, cm, y.packed, y.nwords);
  KLAUNCH_CHECK();
  const KeyGen gen{y.packed, n, 0, 8, 8};
  GSA_TRY(cudaMemsetAsync(y.ghist, 0, MAX_PASSES * RADIX * sizeof(u32), st));
  GSA_TRY(cudaMemsetAsync(y.skip_mask, 0, sizeof(u32), st));
  const u32 hist_blocks = (u32)std::min<u64>((u64)sms * 4, std::max<u64>(1, div_up(n, HIST_THREADS)));
  k_hist0<HIST_THREADS><<<hist_blocks, HIST_THREADS, 0, st>>>(gen, 1, y.ghist);
  KLAUNCH_CHECK();
  k_scan_hist<<<1, RADIX, 0, st>>>(y.ghist, y.bin_base, n, y.skip_mask);
  KLAUNCH_CHECK();
  const u32 tiles = (u32)div_up(n, PASS_TILE);
  GSA_TRY(cudaMemsetAsync(y.pass_status, 0, (256 + (size_t)tiles * RADIX) * sizeof(u32), st));
  PassArgs a;
  a.keys_in = nullptr; a.vals_in = nullptr;
  a.keys_out = y.keys_out; a.vals_out = d_order;
  a.n = n; a.shift = 0;
  a.bin_base = y.bin_base;
  a.counter = y.pass_status; a.status = y.pass_status + 256; a.pf_dist = 0;
  a.gen = gen;
  k_radix_pass<PASS_THREADS, PASS_IPT, true><<<tiles, PASS_THREADS, smem, st>>>(a);
  KLAUNCH_CHECK();
  if (d_counts) GSA_TRY(cudaMemcpyAsync(d_counts, y.ghist, RADIX * sizeof(u32), cudaMemcpyDeviceToDevice, st));
  return GSA_OK;
}

}  // namespace gsa
