// local.cuh -- per-bucket finish: one CTA owns one bucket (or one chunk of an oversized bucket) in shared memory.
//   1. coalesced load, 2. stable LSD radix sort of an index array over the key bits the partition levels have not
//   consumed (10-bit digits, warp-private match ranking, passes whose digit is constant are skipped),
//   3. parallel run detection (head flags + block-scan compaction), 4. consumer.
// Consumers:
//   kCountEmit  : runs of equal keys -> multiplicity, --min-count filter, edge records (KmerCounter::Lv2Postprocess)
//   kCountPairs : a chunk of an oversized bucket -> every distinct (key, count) pair (high-depth keys, e.g. a mitogenome
//                 at 10^4 x, make buckets that no shared memory holds; their chunks are combined first)
//   kCountMerge : the pairs of all chunks of one bucket -> summed multiplicities -> filter -> edge records
//   kSortOnly   : write the bucket back fully sorted (recursive fallback for buckets with too many DISTINCT keys)
//   kSdbgEmit   : BOSS emission per (k-1)-prefix group (SeqToSdbg::Lv2Postprocess), threads own runs of groups
// k_serial runs the same walkers single-threaded over an already sorted global range (last-resort path).
#pragma once
#include "common.cuh"

namespace mf {

enum LocalMode { kCountEmit = 0, kSortOnly = 1, kSdbgEmit = 3, kCountPairs = 4, kCountMerge = 5 };

struct WorkItem {
  int64_t start;   // first record (kCountMerge: first piece index)
  int32_t n;       // records (kCountMerge: number of pieces)
  int32_t slot;    // bucket slot the result belongs to
};

struct LocalArgs {
  const uint32_t *in;        // records
  uint32_t *out;             // kSortOnly: sorted records (a different buffer, same offsets)
  const int64_t *bkt_start;  // [nslots]
  const int64_t *bkt_size;   // [nslots]
  const WorkItem *work;      // [grid] explicit work items, or null: blockIdx.x is the slot
  int bit_off;               // leading bits every record of a bucket shares
  int sort_bits;             // bits that take part in the order (count: 2(k+1); sdbg: all 32 W)
  int cap;                   // max records per CTA (shared memory capacity), even
  int k;
  // --- count ---
  int min_count;
  int words_edge;
  uint32_t *arena;                       // edge records, words_edge words each
  unsigned long long *arena_cursor;      // in records
  unsigned long long arena_cap;          // in records
  int64_t *desc_off;                     // [nslots] arena offset (records)
  int64_t *desc_cnt;                     // [nslots]
  unsigned long long *counting;          // [65536] distinct-edge multiplicity histogram (may be null)
  // --- pairs of oversized buckets ---
  uint32_t *pair_arena;                  // (W+1)-word records: key, count
  unsigned long long *pair_cursor;
  unsigned long long pair_cap;
  int64_t *piece_off;                    // [n_pieces] pair arena offset of each chunk's pairs
  int32_t *piece_cnt;                    // [n_pieces]
  // --- fallback bookkeeping ---
  int32_t *bail_list;                    // slots this launch could not take
  int *bail_count;                       // flags[0]
  int *overflow_flag;                    // flags[1] = edge arena overflow, flags[2] = pair arena overflow
  // --- sdbg ---
  int tip_mode;                          // 0 = seq2sdbg raw tip label, 1 = read2sdbg stage-2 tip-label layout
  int words_tip;
  int64_t *sd_items;                     // [nslots] emitted items
  int64_t *sd_tips;                      // [nslots]
  int64_t *sd_large;                     // [nslots]
  unsigned long long *bucket_stats;      // [65536][3] items / tips / large per megahit bucket (count phase)
  int64_t *sd_item_off;                  // [nslots] arena offset of the bucket's items
  int64_t *sd_tip_off;                   // [nslots] arena offset of the bucket's tip labels
  uint32_t *sd_rec;                      // item arena: w | last<<4 | tip<<5 | mult<<8
  uint32_t *sd_labels;                   // tip-label arena, words_tip per tip
  unsigned long long *sd_cursor;         // [2] item / tip arena cursors
  unsigned long long sd_item_cap, sd_tip_cap;
};

constexpr int kLocalDigitBits = 10;
constexpr int kLocalBins = 1 << kLocalDigitBits;
constexpr int kLocalNT = 256;
constexpr int kHashSlots = 8192;     // open-addressing table of the count consumers (> any cap, so it never fills)
constexpr int kSolidMax = 2048;      // distinct solid keys a bucket may hold on the fast path
constexpr uint32_t kEmptySlot = 0xffffffffu;

__host__ __device__ inline bool mode_is_hash(int mode) { return mode == 0 /*kCountEmit*/ || mode == 4 /*kCountPairs*/ || mode == 5 /*kCountMerge*/; }

// shared memory (uint32 words)
//  LSD family : rec[cap*WR] | idxA,idxB,rk u16[cap] | aux u32[2NT+2] | whist u16[NWARP][1024] | bins u32[1025] | scratch[48]
//  hash family: rec[cap*WR] | table u32[8192] (| tcnt u32[8192] if weighted) | sidx,permA,permB,rk u16[2048] | scnt u32[2048]
//               | small u32[64] | strip u32[NT+1] | scratch[48]          (whist/bins of the solid-key sort alias the table)
inline size_t local_fixed_words(bool hash, bool weighted) {
  if (hash) return (size_t)kHashSlots * (weighted ? 2 : 1) + 4 * (kSolidMax / 2) + kSolidMax + 64 + (kLocalNT + 1) + 48;
  return (size_t)(2 * kLocalNT + 2) + (size_t)(kLocalNT / 32) * kLocalBins / 2 + kLocalBins + 1 + 48;
}
inline size_t local_smem_bytes(int WR, int cap, bool hash, bool weighted) {
  size_t words = (size_t)cap * WR + local_fixed_words(hash, weighted) + (hash ? 0 : 3 * ((size_t)cap / 2));
  return words * 4;
}
// records per CTA such that two CTAs fit an SM (110 KB each); the rare pair / merge launches of oversized buckets take a
// whole SM (kLocalBigSmem) so that far fewer of them fall through to the sorted-run path
constexpr int kLocalBigSmem = 220 * 1024;
inline int local_cap(int WR, bool hash, bool weighted, int smem_budget = 110 * 1024) {
  const int budget = smem_budget - (int)local_fixed_words(hash, weighted) * 4;
  int cap = budget / (WR * 4 + (hash ? 0 : 6));
  cap &= ~1;
  if (hash && cap > kHashSlots - 256) cap = kHashSlots - 256;   // the table must keep free slots
  return cap > 65534 ? 65534 : cap;
}

// ---- sdbg item fields (flag bit 19, b bits 16..18, 65535-mult bits 0..15 of the last word;
//      SeqToSdbg::Lv2ExtractSubString / Extract_a / Extract_b / ExtractCounting)
template <int W>
__device__ __forceinline__ int item_a(const uint32_t *it, int k) {
  if (!((it[W - 1] >> 19) & 1)) return kSentinel;
  int which = (k - 1) >> 4, idx = (k - 1) & 15;
  return (it[which] >> ((15 - idx) * 2)) & 3;
}
template <int W>
__device__ __forceinline__ int item_b(const uint32_t *it) { return (it[W - 1] >> 16) & 7; }
template <int W>
__device__ __forceinline__ bool item_diff_km1(const uint32_t *x, const uint32_t *y, int k) {
  int chars_in_last = (k - 1) & 15, full = (k - 1) >> 4;
  if (chars_in_last > 0) {
    int sh = (16 - chars_in_last) * 2;
    if ((x[full] >> sh) != (y[full] >> sh)) return true;
  }
  for (int i = full - 1; i >= 0; --i)
    if (x[i] != y[i]) return true;
  return false;
}

// compare two records on their leading `bits` bits only
template <int W>
__device__ __forceinline__ int cmp_rec_bits(const uint32_t *a, const uint32_t *b, int bits) {
  const int full = bits >> 5, rem = bits & 31;
  for (int i = 0; i < full; ++i) {
    const uint32_t x = a[i], y = b[i];
    if (x != y) return x < y ? -1 : 1;
  }
  if (rem && full < W) {
    const uint32_t m = 0xffffffffu << (32 - rem);
    const uint32_t x = a[full] & m, y = b[full] & m;
    if (x != y) return x < y ? -1 : 1;
  }
  return 0;
}

// accessors over the sorted order of a range (WR = record stride)
template <int WR>
struct SmemAcc {
  const uint32_t *rec;
  const uint16_t *idx;
  __device__ __forceinline__ const uint32_t *operator()(int i) const { return rec + (size_t)idx[i] * WR; }
};
template <int WR>
struct GmemAcc {
  const uint32_t *base;
  __device__ __forceinline__ const uint32_t *operator()(int64_t i) const { return base + i * WR; }
};

struct SdbgTally {
  uint32_t items, tips, large;
};
// SeqToSdbg::Lv2Postprocess over sorted items [b, e) that form whole (k-1)-prefix groups;
// WRITING=false tallies, true writes at rec_out / lab_out.
template <int W, bool WRITING, class Acc, class I>
__device__ __forceinline__ SdbgTally sdbg_walk(const Acc &acc, I b, I e, int k, int Wt, int tip_mode,
                                               uint32_t *rec_out, uint32_t *lab_out,
                                               unsigned long long *bucket_stats = nullptr) {
  SdbgTally t = {0, 0, 0};
  for (I gs = b, ge; gs < e; gs = ge) {
    const uint32_t *g0 = acc(gs);
    ge = gs + 1;
    while (ge < e && !item_diff_km1<W>(g0, acc(ge), k)) ++ge;
    int has_solid_a = 0, has_solid_b = 0, outputed_b = 0;
    I last_a0 = -1, last_a1 = -1, last_a2 = -1, last_a3 = -1;
    for (I i = gs; i < ge; ++i) {
      const uint32_t *it = acc(i);
      const int ca = item_a<W>(it, k), cb = item_b<W>(it);
      if (ca != kSentinel && cb != kSentinel) { has_solid_a |= 1 << ca; has_solid_b |= 1 << cb; }
      if (ca != kSentinel && (cb != kSentinel || !(has_solid_a & (1 << ca)))) {
        if (ca == 0) last_a0 = i; else if (ca == 1) last_a1 = i; else if (ca == 2) last_a2 = i; else last_a3 = i;
      }
    }
    for (I i = gs, j; i < ge; i = j) {
      const uint32_t *it = acc(i);
      const int ca = item_a<W>(it, k), cb = item_b<W>(it);
      j = i + 1;
      uint32_t inv_mul = it[W - 1] & 0xffffu;   // 65535 - multiplicity: the smallest wins (megahit sorts it first)
      while (j < ge) {
        const uint32_t *nx = acc(j);
        if (item_a<W>(nx, k) != ca || item_b<W>(nx) != cb) break;
        inv_mul = min(inv_mul, nx[W - 1] & 0xffffu);
        ++j;
      }
      int is_dollar = 0;
      if (ca == kSentinel) {
        if (has_solid_b & (1 << cb)) continue;
        is_dollar = 1;
      }
      if (cb == kSentinel) {
        if (has_solid_a & (1 << ca)) continue;
      }
      const int w = (cb == kSentinel) ? 0 : ((outputed_b & (1 << cb)) ? cb + 5 : cb + 1);
      outputed_b |= 1 << cb;
      const I la = ca == 0 ? last_a0 : ca == 1 ? last_a1 : ca == 2 ? last_a2 : last_a3;
      const int last = (ca == kSentinel) ? 0 : (la == j - 1 ? 1 : 0);
      const int mul = kMaxMul - (int)inv_mul;
      if (WRITING) {
        rec_out[t.items] = (uint32_t)w | ((uint32_t)last << 4) | ((uint32_t)is_dollar << 5) | ((uint32_t)mul << 8);
        if (is_dollar) {
          uint32_t *lab = lab_out + (size_t)t.tips * Wt;
          for (int q = 0; q < Wt; ++q) lab[q] = it[q];      // raw first words of the item
          if (tip_mode == 1) {
            // read2sdbg stage-2 items carry only flag<<3 | b below the bases (Read2SdbgS2::Lv2ExtractSubString)
            if (W == Wt) lab[Wt - 1] &= 0xfff00000u;
            if ((2 * k + 4 + 31) / 32 == Wt) lab[Wt - 1] |= (uint32_t)cb;
          }
        }
      }
      if (!WRITING && bucket_stats) {
        unsigned long long *bs = bucket_stats + (size_t)(it[0] >> 16) * 3;
        atomicAdd(bs, 1ull);
        if (is_dollar) atomicAdd(bs + 1, 1ull);
        if (mul > 254) atomicAdd(bs + 2, 1ull);
      }
      t.items += 1;
      t.tips += is_dollar;
      t.large += mul > 254;
    }
  }
  return t;
}

// KmerCounter::PackEdge: key words, zero padding, multiplicity in the low 16 bits of the last word
template <int W>
__device__ __forceinline__ void write_edge(uint32_t *dst, const uint32_t *key, int We, uint32_t count) {
  for (int t = 0; t < We; ++t) dst[t] = t < W ? key[t] : 0u;
  dst[We - 1] |= count > (uint32_t)kMaxMul ? (uint32_t)kMaxMul : count;
}

// One stable LSD pass over the index array.  Returns false (and leaves idx_in as the current order) when every record
// has the same digit.  Warps own contiguous blocks of positions, so warp order == position order == stability.
// peer mask of lanes holding the same digit, by ballots (the ADU pipe takes ~3.6 cycles per ballot, ~63 per match.any)
__device__ __forceinline__ unsigned peers_by_ballot(uint32_t d, bool valid, int nbits) {
  unsigned m = __ballot_sync(0xffffffffu, valid);
  for (int b = 0; b < nbits; ++b) {
    const bool bit = (d >> b) & 1;
    const unsigned v = __ballot_sync(0xffffffffu, bit);
    m &= bit ? v : ~v;
  }
  return valid ? m : 0u;
}
template <int W, bool SWAP64>
__device__ __forceinline__ uint32_t lsd_digit(const uint32_t *r, int bit_lo, int nbits) {
  if constexpr (SWAP64) {   // a u64 key stored natively: word 1 is the high half
    const unsigned long long key = ((unsigned long long)r[1] << 32) | r[0];
    return (uint32_t)((key << bit_lo) >> (64 - nbits));
  } else {
    return rec_digit_mem<W>(r, bit_lo, nbits);
  }
}
// `via` (nullable) maps sorted entities to records: record = via[entity]
template <int WR, int W, int NT, bool SWAP64 = false>
__device__ __forceinline__ bool lsd_pass(const uint32_t *rec, const uint16_t *idx_in, uint16_t *idx_out, uint16_t *rk, int n,
                                         int bit_lo, int nbits, uint16_t *whist, uint32_t *bins, uint32_t *scratch,
                                         const uint16_t *via = nullptr) {
  constexpr int NWARP = NT / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = 1 << nbits;
  const int per_warp = (((n + NWARP - 1) / NWARP) + 31) & ~31;
  for (int i = tid; i < NWARP * nb / 2; i += NT) reinterpret_cast<uint32_t *>(whist)[i] = 0;
  if (NWARP * nb < 2 && tid == 0) whist[0] = 0;
  __syncthreads();
  uint16_t *wh = whist + warp * nb;
  const int p0 = warp * per_warp;
  for (int it = 0; it < per_warp; it += 32) {
    const int p = p0 + it + lane;
    const bool valid = p < n;
    uint32_t d = 0;
    if (valid) {
      int r = idx_in ? idx_in[p] : p;
      if (via) r = via[r];
      d = lsd_digit<W, SWAP64>(rec + (size_t)r * WR, bit_lo, nbits);
    }
    const unsigned m = peers_by_ballot(d, valid, nbits);
    const unsigned leader = (unsigned)(__ffs(m) - 1);
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = wh[d];
      wh[d] = (uint16_t)(old + __popc(m));
    }
    old = __shfl_sync(0xffffffffu, old, valid ? leader : 0);
    if (valid) rk[p] = (uint16_t)(old + __popc(m & lanemask_lt()));
    __syncwarp();
  }
  __syncthreads();
  int single = 0;
  for (int b = tid; b < nb; b += NT) {
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
      uint32_t c = whist[w * nb + b];
      whist[w * nb + b] = (uint16_t)acc;
      acc += c;
    }
    bins[b] = acc;
    single |= acc == (uint32_t)n;
  }
  if (tid == 0) bins[nb] = 0;
  single = __syncthreads_or(single);
  if (single) return false;
  block_excl_scan<NT>(bins, nb + 1, scratch);
  for (int it = 0; it < per_warp; it += 32) {
    const int p = p0 + it + lane;
    if (p < n) {
      const int e = idx_in ? idx_in[p] : p;
      const int r = via ? via[e] : e;
      const uint32_t d = lsd_digit<W, SWAP64>(rec + (size_t)r * WR, bit_lo, nbits);
      idx_out[bins[d] + wh[d] + rk[p]] = (uint16_t)e;
    }
  }
  __syncthreads();
  return true;
}

// Block-wide compaction of flagged positions in [0, n): list[j] = j-th flagged position (ascending); returns the count.
// flag(p) is evaluated twice.  Threads own contiguous strips, so the list is ordered.
template <int NT, class F>
__device__ __forceinline__ uint32_t compact_positions(int n, F flag, uint16_t *list, uint32_t *strip_cnt /*[NT+1]*/,
                                                      uint32_t *scratch, uint32_t max_out = 0xffffffffu) {
  const int tid = threadIdx.x;
  const int per = (n + NT - 1) / NT;
  const int b = min(n, tid * per), e = min(n, b + per);
  uint32_t c = 0;
  for (int p = b; p < e; ++p) c += flag(p) ? 1u : 0u;
  strip_cnt[tid] = c;
  if (tid == 0) strip_cnt[NT] = 0;
  __syncthreads();
  const uint32_t total = block_excl_scan<NT>(strip_cnt, NT + 1, scratch);
  if (total <= max_out) {
    uint32_t o = strip_cnt[tid];
    for (int p = b; p < e; ++p)
      if (flag(p)) list[o++] = (uint16_t)p;
  }
  __syncthreads();
  return total;
}

// ---------------------------------------------------------------------------------------------------------
// LSD family: kSortOnly (fallback) and kSdbgEmit.  The whole bucket is sorted (all sort_bits), then consumed.
template <int W, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_local(LocalArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NWARP = NT / 32;
  constexpr int WR = W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cap = a.cap;

  uint32_t *rec = smem;                                                       // [cap*WR]
  uint16_t *idxA = reinterpret_cast<uint16_t *>(rec + (size_t)cap * WR);      // [cap]
  uint16_t *idxB = idxA + cap;                                                // [cap]
  uint16_t *rk = idxB + cap;                                                  // [cap]
  uint32_t *aux = reinterpret_cast<uint32_t *>(rk + cap);                     // [2NT+2]
  uint16_t *whist = reinterpret_cast<uint16_t *>(aux + 2 * NT + 2);           // [NWARP][1024]
  uint32_t *bins = reinterpret_cast<uint32_t *>(whist + NWARP * kLocalBins);  // [1025]
  uint32_t *scratch = bins + kLocalBins + 1;                                  // [34]
  int *s_flag = reinterpret_cast<int *>(scratch + 34);                        // [8]

  int slot;
  int64_t start, n64;
  if (a.work) {
    const WorkItem wi = a.work[blockIdx.x];
    slot = wi.slot;
    start = wi.start;
    n64 = wi.n;
  } else {
    slot = (int)blockIdx.x;
    start = a.bkt_start[slot];
    n64 = a.bkt_size[slot];
  }
  if (n64 == 0) return;
  if (n64 > cap) {                       // does not fit: the fallback path takes it
    if (tid == 0) {
      int p = atomicAdd(a.bail_count, 1);
      a.bail_list[p] = slot;
    }
    return;
  }
  const int n = (int)n64;
  {
    const uint32_t *src = a.in + start * (int64_t)W;
    const int nw = n * W;
    if constexpr (W % 2 == 0) {
      const uint2 *s2 = reinterpret_cast<const uint2 *>(src);
      uint2 *d2 = reinterpret_cast<uint2 *>(rec);
      for (int i = tid; i < nw / 2; i += NT) d2[i] = s2[i];
    } else {
      for (int i = tid; i < nw; i += NT) rec[i] = src[i];
    }
  }
  if (tid < 8) s_flag[tid] = 0;
  __syncthreads();

  const uint16_t *cur = nullptr;   // nullptr == identity order
  bool sorted = false;
  // LSD sort of the index array over bits [bit_off, sort_bits), least significant digit first
  if (!sorted) {
    uint16_t *nxt = idxA;
    int hi = a.sort_bits;
    while (hi > a.bit_off) {
      const int nb = min(kLocalDigitBits, hi - a.bit_off);
      if (lsd_pass<WR, W, NT>(rec, cur, nxt, rk, n, hi - nb, nb, whist, bins, scratch)) {
        cur = nxt;
        nxt = (nxt == idxA) ? idxB : idxA;
      }
      hi -= nb;
    }
    if (!cur) {   // never permuted: materialise the identity so that the consumers can index
      for (int i = tid; i < n; i += NT) idxA[i] = (uint16_t)i;
      cur = idxA;
      __syncthreads();
    }
  }
  uint16_t *list = (cur == idxA) ? idxB : idxA;   // the free index array: group positions
  const SmemAcc<WR> acc{rec, cur};
  uint32_t *strip = bins;   // [NT+1] fits in bins[1025]

  if constexpr (MODE == kSortOnly) {
    uint32_t *dst = a.out + start * (int64_t)W;
    const int nw = n * W;
    for (int x = tid; x < nw; x += NT) {
      int p = x / W, c = x - p * W;
      dst[x] = rec[(size_t)cur[p] * W + c];
    }
    return;
  }

  if constexpr (MODE == kSdbgEmit) {
    // groups of equal (k-1)-prefix; threads own contiguous runs of groups (a group is a handful of items)
    auto is_ghead = [&](int p) { return p == 0 || item_diff_km1<W>(acc(p), acc(p - 1), a.k); };
    const uint32_t ng = compact_positions<NT>(n, is_ghead, list, strip, scratch);
    uint32_t my_items = 0, my_tips = 0, my_large = 0;
    const int gper = ((int)ng + NT - 1) / NT;
    const int gb = min((int)ng, tid * gper), ge = min((int)ng, gb + gper);
    for (int g = gb; g < ge; ++g) {
      const int p0 = list[g], p1 = (g + 1 < (int)ng) ? (int)list[g + 1] : n;
      SdbgTally t = sdbg_walk<W, false>(acc, p0, p1, a.k, a.words_tip, a.tip_mode, nullptr, nullptr, a.bucket_stats);
      my_items += t.items;
      my_tips += t.tips;
      my_large += t.large;
    }
    uint32_t *s_items = aux, *s_tips = aux + NT + 1;
    s_items[tid] = my_items;
    s_tips[tid] = my_tips;
    if (tid == 0) { s_items[NT] = 0; s_tips[NT] = 0; }
    __syncthreads();
    const uint32_t tot_items = block_excl_scan<NT>(s_items, NT + 1, scratch);
    const uint32_t tot_tips = block_excl_scan<NT>(s_tips, NT + 1, scratch);
#pragma unroll
    for (int o = 16; o; o >>= 1) my_large += __shfl_xor_sync(0xffffffffu, my_large, o);
    if (lane == 0) scratch[warp] = my_large;
    __syncthreads();
    if (tid == 0) {
      uint32_t L = 0;
      for (int w = 0; w < NWARP; ++w) L += scratch[w];
      const unsigned long long ib = atomicAdd(a.sd_cursor, (unsigned long long)tot_items);
      const unsigned long long tb = atomicAdd(a.sd_cursor + 1, (unsigned long long)tot_tips);
      const int ok = ib + tot_items <= a.sd_item_cap && tb + tot_tips <= a.sd_tip_cap;
      if (!ok) atomicExch(a.overflow_flag, 1);
      a.sd_item_off[slot] = (int64_t)ib;
      a.sd_tip_off[slot] = (int64_t)tb;
      a.sd_items[slot] = ok ? (int64_t)tot_items : 0;
      a.sd_tips[slot] = ok ? (int64_t)tot_tips : 0;
      a.sd_large[slot] = ok ? (int64_t)L : 0;
      s_flag[1] = ok;
      s_flag[2] = (int)(uint32_t)ib; s_flag[3] = (int)(uint32_t)(ib >> 32);
      s_flag[4] = (int)(uint32_t)tb; s_flag[5] = (int)(uint32_t)(tb >> 32);
    }
    __syncthreads();
    if (!s_flag[1]) return;
    const int64_t item_base = (int64_t)(((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2]) + s_items[tid];
    const int64_t tip_base = (int64_t)(((unsigned long long)(uint32_t)s_flag[5] << 32) | (uint32_t)s_flag[4]) + s_tips[tid];
    uint32_t oi = 0, ot = 0;
    for (int g = gb; g < ge; ++g) {
      const int p0 = list[g], p1 = (g + 1 < (int)ng) ? (int)list[g + 1] : n;
      SdbgTally t = sdbg_walk<W, true>(acc, p0, p1, a.k, a.words_tip, a.tip_mode, a.sd_rec + item_base + oi,
                                       a.sd_labels + (tip_base + ot) * (int64_t)a.words_tip);
      oi += t.items;
      ot += t.tips;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// hash family: count consumers.  Equal keys are grouped by an open-addressing table in shared memory (one CAS to claim or
// find the slot, one atomicAdd for the count); only the distinct SOLID keys -- a few per cent of the bucket at
// assembly depths -- are then sorted (ballot-ranked stable LSD over the unconsumed key bits) and emitted.
template <int W>
__device__ __forceinline__ uint32_t hash_key(const uint32_t *k) {
  uint32_t h = 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    h ^= k[i];
    h *= 0x85ebca6bu;
    h ^= h >> 13;
  }
  h *= 0xc2b2ae35u;
  return h ^ (h >> 16);
}

template <int W, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_count(LocalArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NWARP = NT / 32;
  constexpr bool WEIGHTED = MODE == kCountMerge;
  constexpr int WR = WEIGHTED ? W + 1 : W;
  const int tid = threadIdx.x;
  const int cap = a.cap;

  uint32_t *rec = smem;                                                  // [cap*WR]
  uint32_t *table = rec + (size_t)cap * WR;                              // [8192] owner<<16 | count   (weighted: owner)
  uint32_t *tcnt = table + kHashSlots;                                   // [8192] weighted only
  uint16_t *sidx = reinterpret_cast<uint16_t *>(WEIGHTED ? tcnt + kHashSlots : tcnt);   // [2048] solid key -> record
  uint16_t *permA = sidx + kSolidMax, *permB = permA + kSolidMax, *rk = permB + kSolidMax;
  uint32_t *scnt = reinterpret_cast<uint32_t *>(rk + kSolidMax);         // [2048] solid key -> multiplicity
  uint32_t *s_small = scnt + kSolidMax;                                  // [64]
  uint32_t *strip = s_small + 64;                                        // [NT+1]
  uint32_t *scratch = strip + NT + 1;                                    // [34]
  int *s_flag = reinterpret_cast<int *>(scratch + 34);                   // [8]
  // the solid-key sort reuses the table once it has been drained
  uint16_t *whist = reinterpret_cast<uint16_t *>(table);                 // [NWARP][1024] = 16 KB
  uint32_t *bins = table + NWARP * kLocalBins / 2;                       // [1025]

  int slot;
  int64_t start, n64;
  if (a.work) {
    const WorkItem wi = a.work[blockIdx.x];
    slot = wi.slot;
    start = wi.start;
    n64 = wi.n;
  } else {
    slot = (int)blockIdx.x;
    start = a.bkt_start[slot];
    n64 = a.bkt_size[slot];
  }
  if (n64 == 0) return;
  int n;
  auto bail = [&]() {
    if (tid == 0) {
      int p = atomicAdd(a.bail_count, 1);
      a.bail_list[p] = slot;
    }
  };
  // ---- 1. load
  if constexpr (MODE == kCountMerge) {
    const int first = (int)start, np = (int)n64;   // pieces: every chunk's (key, count) pairs
    int total = 0;
    for (int q = 0; q < np; ++q) total += a.piece_cnt[first + q];
    if (total > cap) { bail(); return; }
    int o = 0;
    for (int q = 0; q < np; ++q) {
      const int c = a.piece_cnt[first + q];
      const uint32_t *src = a.pair_arena + a.piece_off[first + q] * (int64_t)WR;
      for (int i = tid; i < c * WR; i += NT) rec[(size_t)o * WR + i] = src[i];
      o += c;
    }
    n = total;
    if (n == 0) {
      if (tid == 0) { a.desc_off[slot] = 0; a.desc_cnt[slot] = 0; }
      return;
    }
  } else {
    if (n64 > cap) { bail(); return; }
    n = (int)n64;
    const uint32_t *src = a.in + start * (int64_t)W;
    const int nw = n * W;
    if constexpr (W % 2 == 0) {
      const uint2 *s2 = reinterpret_cast<const uint2 *>(src);
      uint2 *d2 = reinterpret_cast<uint2 *>(rec);
      for (int i = tid; i < nw / 2; i += NT) d2[i] = s2[i];
    } else {
      for (int i = tid; i < nw; i += NT) rec[i] = src[i];
    }
  }
  // table size follows the bucket: a power of two >= 2n keeps probes short, and every block-wide sweep is over `ts` slots
  int ts = 256;
  while (ts < 2 * n && ts < kHashSlots) ts <<= 1;
  const uint32_t tmask = (uint32_t)ts - 1;
  for (int i = tid; i < ts; i += NT) {
    table[i] = kEmptySlot;
    if constexpr (WEIGHTED) tcnt[i] = 0u;
  }
  if (tid < 64) s_small[tid] = 0;
  if (tid < 8) s_flag[tid] = 0;
  __syncthreads();

  // ---- 2. group equal keys
  for (int i = tid; i < n; i += NT) {
    const uint32_t *key = rec + (size_t)i * WR;
    uint32_t h = hash_key<W>(key) & tmask;
    for (;;) {
      uint32_t cur = *reinterpret_cast<volatile uint32_t *>(table + h);
      if (cur == kEmptySlot) {
        const uint32_t mine = WEIGHTED ? (uint32_t)i : (((uint32_t)i << 16) | 1u);
        cur = atomicCAS(table + h, kEmptySlot, mine);
        if (cur == kEmptySlot) {
          if constexpr (WEIGHTED) atomicAdd(tcnt + h, key[W]);
          break;
        }
      }
      const uint32_t owner = WEIGHTED ? cur : (cur >> 16);
      if (cmp_rec<W>(rec + (size_t)owner * WR, key) == 0) {
        if constexpr (WEIGHTED) atomicAdd(tcnt + h, key[W]);
        else atomicAdd(table + h, 1u);
        break;
      }
      h = (h + 1) & tmask;
    }
  }
  __syncthreads();
  auto slot_count = [&](int h) -> uint32_t { return WEIGHTED ? tcnt[h] : (table[h] & 0xffffu); };
  auto slot_owner = [&](int h) -> uint32_t { return WEIGHTED ? table[h] : (table[h] >> 16); };

  if constexpr (MODE == kCountPairs) {
    // every distinct key of this chunk with its count -> pair arena (unsorted: the merge groups them again)
    const int PER = ts / NT;   // ts >= 256 == NT
    uint32_t c = 0;
    for (int h = tid * PER; h < (tid + 1) * PER; ++h) c += table[h] != kEmptySlot;
    strip[tid] = c;
    if (tid == 0) strip[NT] = 0;
    __syncthreads();
    const uint32_t nd = block_excl_scan<NT>(strip, NT + 1, scratch);
    if (tid == 0) {
      unsigned long long base = atomicAdd(a.pair_cursor, (unsigned long long)nd);
      int ok = base + nd <= a.pair_cap;
      if (!ok) atomicExch(a.overflow_flag + 1, 1);
      a.piece_off[blockIdx.x] = (int64_t)base;
      a.piece_cnt[blockIdx.x] = ok ? (int32_t)nd : 0;
      s_flag[1] = ok;
      s_flag[2] = (int)(uint32_t)base;
      s_flag[3] = (int)(uint32_t)(base >> 32);
    }
    __syncthreads();
    if (!s_flag[1]) return;
    const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
    uint32_t o = strip[tid];
    for (int h = tid * PER; h < (tid + 1) * PER; ++h) {
      if (table[h] == kEmptySlot) continue;
      const uint32_t *key = rec + (size_t)slot_owner(h) * WR;
      uint32_t *dst = a.pair_arena + (base + o) * (unsigned long long)(W + 1);
#pragma unroll
      for (int t = 0; t < W; ++t) dst[t] = key[t];
      dst[W] = slot_count(h);
      ++o;
    }
    return;
  } else {
    // ---- 3. solid keys: compact (bail before any side effect if they exceed the fast path)
    auto is_solid = [&](int h) { return table[h] != kEmptySlot && slot_count(h) >= (uint32_t)a.min_count; };
    const uint32_t ns = compact_positions<NT>(ts, is_solid, permB /* slot list */, strip, scratch, kSolidMax);
    if (ns > (uint32_t)kSolidMax) { bail(); return; }
    // distinct-edge multiplicity histogram (<prefix>.counting), hot small counts stay in shared memory
    if (a.counting) {
      for (int h = tid; h < ts; h += NT) {
        if (table[h] == kEmptySlot) continue;
        const uint32_t c = min(slot_count(h), (uint32_t)kMaxMul);
        if (c < 64) atomicAdd(s_small + c, 1u);
        else atomicAdd(a.counting + c, 1ull);
      }
      __syncthreads();
      if (tid < 64 && s_small[tid]) atomicAdd(a.counting + tid, (unsigned long long)s_small[tid]);
    }
    for (uint32_t q = tid; q < ns; q += NT) {
      const int h = permB[q];
      sidx[q] = (uint16_t)slot_owner(h);
      scnt[q] = slot_count(h);
    }
    if (tid == 0) {
      unsigned long long base = atomicAdd(a.arena_cursor, (unsigned long long)ns);
      int ok = base + ns <= a.arena_cap;
      if (!ok) atomicExch(a.overflow_flag, 1);
      a.desc_off[slot] = (int64_t)base;
      a.desc_cnt[slot] = ok ? (int64_t)ns : 0;
      s_flag[1] = ok;
      s_flag[2] = (int)(uint32_t)base;
      s_flag[3] = (int)(uint32_t)(base >> 32);
    }
    __syncthreads();   // the table is dead from here on (whist / bins alias it)
    if (!s_flag[1] || ns == 0) return;
    // ---- 4. order the solid keys.  A handful (the usual case): rank by all-pairs comparison; many: stable LSD passes.
    const uint16_t *cur = nullptr;
    if (ns > 1 && ns <= 128) {
      for (uint32_t q = tid; q < ns; q += NT) {   // keys are distinct: rank = number of smaller keys
        const uint32_t *kq = rec + (size_t)sidx[q] * WR;
        uint32_t r = 0;
        for (uint32_t o = 0; o < ns; ++o) r += cmp_rec<W>(rec + (size_t)sidx[o] * WR, kq) < 0;
        permA[r] = (uint16_t)q;
      }
      __syncthreads();
      cur = permA;
    } else if (ns > 128) {
      uint16_t *nxt = permA;
      int hi = a.sort_bits;
      while (hi > a.bit_off) {
        const int nb = min(kLocalDigitBits, hi - a.bit_off);
        if (lsd_pass<WR, W, NT>(rec, cur, nxt, rk, (int)ns, hi - nb, nb, whist, bins, scratch, sidx)) {
          cur = nxt;
          nxt = (nxt == permA) ? permB : permA;
        }
        hi -= nb;
      }
    }
    const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
    const int We = a.words_edge;
    for (uint32_t q = tid; q < ns; q += NT) {
      const uint32_t e = cur ? cur[q] : q;
      write_edge<W>(a.arena + (base + q) * (unsigned long long)We, rec + (size_t)sidx[e] * WR, We, scnt[e]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fast count path for keys of at most 64 bits (k <= 31): the bucket is STREAMED from global memory into a shared hash
// table that stores the keys themselves (64-bit CAS) and 32-bit counts, so a bucket of any size fits as long as its
// DISTINCT keys do -- a mitochondrial (k+1)-mer seen 10^4 times is one slot.  A thread whose atomicAdd carries a count
// across --min-count appends the slot to the solid list; only those keys are ordered and emitted.
constexpr int kFastSlots = 4096;
constexpr int kFastProbeLimit = 96;
constexpr int kFastNT = 512;
constexpr int kFastSolidMax = 1024;   // distinct solid keys of one bucket on the fast path
constexpr unsigned long long kEmptyKey = 0xffffffffffffffffull;   // never a canonical key

inline size_t fast_smem_bytes() {
  // tkeys u64[4096] | tcnt u32[4096] | sidx,permA,permB,rk u16[1024] | scnt u32[1024] | bins u32[1025] | small u32[64] | scratch
  // | whist u16[NWARP][1024]
  return (size_t)kFastSlots * 12 + 4 * kFastSolidMax * 2 + kFastSolidMax * 4 + (kLocalBins + 1) * 4 + 64 * 4 + 64 * 4 +
         (size_t)(kFastNT / 32) * kLocalBins * 2;
}

template <int W, int NT>
__global__ void __launch_bounds__(NT) k_count_fast(LocalArgs a) {
  static_assert(W <= 2, "keys must fit 64 bits");
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NWARP = NT / 32;
  const int tid = threadIdx.x;
  unsigned long long *tkeys = reinterpret_cast<unsigned long long *>(smem);            // [4096]
  uint32_t *tcnt = smem + 2 * kFastSlots;                                              // [4096]
  uint16_t *sidx = reinterpret_cast<uint16_t *>(tcnt + kFastSlots);                    // [2048] solid -> slot
  uint16_t *permA = sidx + kFastSolidMax, *permB = permA + kFastSolidMax, *rk = permB + kFastSolidMax;
  uint32_t *scnt = reinterpret_cast<uint32_t *>(rk + kFastSolidMax);                       // [2048]
  uint32_t *bins = scnt + kFastSolidMax;                                                   // [1025]
  uint32_t *s_small = bins + kLocalBins + 1;                                           // [64]
  uint32_t *scratch = s_small + 64;                                                    // [34]
  int *s_flag = reinterpret_cast<int *>(scratch + 34);                                 // [8]: 0 bail, 1 ok, 2..3 base, 4 ns
  uint16_t *whist = reinterpret_cast<uint16_t *>(s_small + 128);                       // [NWARP][1024] (many-solid-keys sort)

  const int slot = a.work ? a.work[blockIdx.x].slot : (int)blockIdx.x;
  const int64_t start = a.bkt_start[slot];
  const int64_t n = a.bkt_size[slot];
  if (n == 0) return;
  // the first keys of every thread are requested before the table is cleared, so their DRAM latency hides behind it
  constexpr int PF = 8;
  const uint2 *src2 = reinterpret_cast<const uint2 *>(a.in) + start;
  uint2 pre[PF];
  if constexpr (W == 2) {
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      const int64_t i = (int64_t)tid + (int64_t)q * NT;
      pre[q] = i < n ? src2[i] : make_uint2(0u, 0u);
    }
  }
  // table size follows the bucket (distinct keys are a fraction of it): fewer slots to clear and to sweep
  int ts_log = 9;
  while ((1 << ts_log) < n && ts_log < 12) ++ts_log;
  const int ts = 1 << ts_log;
  for (int i = tid; i < ts; i += NT) {
    tkeys[i] = kEmptyKey;
    tcnt[i] = 0u;
  }
  if (tid < 64) s_small[tid] = 0;
  if (tid < 8) s_flag[tid] = 0;
  __syncthreads();

  // ---- 1. stream the bucket through the table
  const uint32_t m = (uint32_t)a.min_count;
  auto insert = [&](unsigned long long key) {
    uint32_t h = (uint32_t)((key * 0x9e3779b97f4a7c15ull) >> (64 - ts_log));
    for (int probe = 0; probe < kFastProbeLimit; ++probe) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tkeys + h);
      if (cur == kEmptyKey) cur = atomicCAS(tkeys + h, kEmptyKey, key);
      if (cur == kEmptyKey || cur == key) {
        const uint32_t old = atomicAdd(tcnt + h, 1u);
        if (old + 1 == m) {
          const int q = atomicAdd(s_flag + 4, 1);
          if (q < kFastSolidMax) sidx[q] = (uint16_t)h;
        }
        return;
      }
      h = (h + 1) & (ts - 1);
    }
    s_flag[0] = 1;   // table too crowded: this bucket takes the general path
  };
  if constexpr (W == 2) {
    int64_t i = (int64_t)tid + (int64_t)PF * NT;
    for (int q = 0; q < PF; ++q)
      if ((int64_t)tid + (int64_t)q * NT < n) insert(((unsigned long long)pre[q].x << 32) | pre[q].y);
    for (; i + 3 * NT < n; i += 4 * NT) {
      const uint2 v0 = src2[i], v1 = src2[i + NT], v2 = src2[i + 2 * NT], v3 = src2[i + 3 * NT];
      insert(((unsigned long long)v0.x << 32) | v0.y);
      insert(((unsigned long long)v1.x << 32) | v1.y);
      insert(((unsigned long long)v2.x << 32) | v2.y);
      insert(((unsigned long long)v3.x << 32) | v3.y);
    }
    for (; i < n; i += NT) {
      const uint2 v = src2[i];
      insert(((unsigned long long)v.x << 32) | v.y);
    }
  } else {
    const uint32_t *src = a.in + start;
    for (int64_t i = tid; i < n; i += NT) insert((unsigned long long)src[i] << 32);
  }
  __syncthreads();
  const int ns_raw = s_flag[4];
  if (s_flag[0] || ns_raw > kFastSolidMax) {
    if (tid == 0) {
      int p = atomicAdd(a.bail_count, 1);
      a.bail_list[p] = slot;
    }
    return;
  }
  const uint32_t ns = (uint32_t)ns_raw;
  // ---- 2. distinct-edge multiplicity histogram (<prefix>.counting)
  if (a.counting) {
    for (int h = tid; h < ts; h += NT) {
      if (tkeys[h] == kEmptyKey) continue;
      const uint32_t c = min(tcnt[h], (uint32_t)kMaxMul);
      if (c < 64) atomicAdd(s_small + c, 1u);
      else atomicAdd(a.counting + c, 1ull);
    }
    __syncthreads();
    if (tid < 64 && s_small[tid]) atomicAdd(a.counting + tid, (unsigned long long)s_small[tid]);
  }
  for (uint32_t q = tid; q < ns; q += NT) scnt[q] = tcnt[sidx[q]];
  if (tid == 0) {
    unsigned long long base = atomicAdd(a.arena_cursor, (unsigned long long)ns);
    int ok = base + ns <= a.arena_cap;
    if (!ok) atomicExch(a.overflow_flag, 1);
    a.desc_off[slot] = (int64_t)base;
    a.desc_cnt[slot] = ok ? (int64_t)ns : 0;
    s_flag[1] = ok;
    s_flag[2] = (int)(uint32_t)base;
    s_flag[3] = (int)(uint32_t)(base >> 32);
  }
  __syncthreads();   // tcnt is dead from here on (the LSD scratch aliases it)
  if (!s_flag[1] || ns == 0) return;
  // ---- 3. order the solid keys (distinct): all-pairs rank for a handful, stable LSD passes otherwise
  const uint16_t *cur = nullptr;
  if (ns > 1 && ns <= 160) {
    uint32_t *rank = bins;   // [ns] <= 1025
    for (uint32_t q = tid; q < ns; q += NT) rank[q] = 0;
    __syncthreads();
    for (uint32_t x = tid; x < ns * ns; x += NT) {   // every pair once, spread over the whole CTA
      const uint32_t q = x / ns, o = x - q * ns;
      if (tkeys[sidx[o]] < tkeys[sidx[q]]) atomicAdd(rank + q, 1u);
    }
    __syncthreads();
    for (uint32_t q = tid; q < ns; q += NT) permA[rank[q]] = (uint16_t)q;
    __syncthreads();
    cur = permA;
  } else if (ns > 160) {
    // keys live in the table as (hi, lo) words == little-endian u64: present them big-endian to the digit reader
    const uint32_t *rec = reinterpret_cast<const uint32_t *>(tkeys);
    uint16_t *nxt = permA;
    int hi = a.sort_bits;
    while (hi > a.bit_off) {
      const int nb = min(kLocalDigitBits, hi - a.bit_off);
      if (lsd_pass<2, 2, NT, true>(rec, cur, nxt, rk, (int)ns, hi - nb, nb, whist, bins, scratch, sidx)) {
        cur = nxt;
        nxt = (nxt == permA) ? permB : permA;
      }
      hi -= nb;
    }
  }
  const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
  const int We = a.words_edge;
  for (uint32_t q = tid; q < ns; q += NT) {
    const uint32_t e = cur ? cur[q] : q;
    const unsigned long long key = tkeys[sidx[e]];
    const uint32_t kw[2] = {(uint32_t)(key >> 32), (uint32_t)key};
    write_edge<W>(a.arena + (base + q) * (unsigned long long)We, kw, We, scnt[e]);
  }
}


// Last-resort count finish, parallel: the range is fully sorted in global memory; one CTA per range finds the runs of equal
// keys with head / tail flags (a run's length = tail position - position of the last head at or before it, a block-wide
// max-scan with a carry between tiles), in two sweeps: total of the solid runs -> one arena reservation -> records.
// Replaces a one-thread-per-range walker that took 100 ms for a 100 K-key range.
template <int W, int NT>
__global__ void __launch_bounds__(NT) k_sorted_runs(LocalArgs a, int nwork) {
  __shared__ long long s_scan[NT];
  __shared__ uint32_t s_cnt[NT + 1];
  __shared__ uint32_t s_scratch[40];
  __shared__ long long s_carry_head;
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_emitted;
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  const int slot = a.work[blockIdx.x].slot;
  const int64_t n = a.bkt_size[slot];
  const uint32_t *keys = a.in + a.bkt_start[slot] * (int64_t)W;
  if (n == 0) return;
  auto differs = [&](int64_t i, int64_t j) {
    const uint32_t *x = keys + i * W, *y = keys + j * W;
    bool d = false;
#pragma unroll
    for (int q = 0; q < W; ++q) d = d || (x[q] != y[q]);
    return d;
  };
  for (int sweep = 0; sweep < 2; ++sweep) {
    if (tid == 0) { s_carry_head = 0; s_emitted = 0; }
    __syncthreads();
    for (int64_t i0 = 0; i0 < n; i0 += NT) {
      const int64_t i = i0 + tid;
      const bool in_range = i < n;
      const bool head = in_range && (i == 0 || differs(i, i - 1));
      const bool tail = in_range && (i == n - 1 || differs(i, i + 1));
      // position of the last head at or before i: inclusive max-scan over the tile, seeded with the carry
      s_scan[tid] = head ? (long long)i : -1;
      __syncthreads();
      for (int o = 1; o < NT; o <<= 1) {
        const long long v = tid >= o ? s_scan[tid - o] : -1;
        __syncthreads();
        if (v > s_scan[tid]) s_scan[tid] = v;
        __syncthreads();
      }
      const long long hp = s_scan[tid] >= 0 ? s_scan[tid] : s_carry_head;
      const long long len = tail ? (long long)i - hp + 1 : 0;
      const bool solid = tail && len >= a.min_count;
      s_cnt[tid] = solid ? 1u : 0u;
      if (tid == 0) s_cnt[NT] = 0;
      __syncthreads();
      const uint32_t tile_total = block_excl_scan<NT>(s_cnt, NT + 1, s_scratch);
      if (tail) {
        const uint32_t c = len > kMaxMul ? (uint32_t)kMaxMul : (uint32_t)len;
        if (sweep == 1) {
          if (a.counting) atomicAdd(a.counting + c, 1ull);
          if (solid && s_ok)
            write_edge<W>(a.arena + (s_base + s_emitted + s_cnt[tid]) * (unsigned long long)a.words_edge, keys + i * W, a.words_edge, c);
        }
      }
      __syncthreads();
      if (tid == 0) {
        const long long last = s_scan[NT - 1];
        if (last >= 0) s_carry_head = last;
        s_emitted += tile_total;
      }
      __syncthreads();
    }
    if (sweep == 0) {
      if (tid == 0) {
        const unsigned long long total = s_emitted;
        const unsigned long long base = atomicAdd(a.arena_cursor, total);
        const int ok = base + total <= a.arena_cap;
        if (!ok) atomicExch(a.overflow_flag, 1);
        a.desc_off[slot] = (int64_t)base;
        a.desc_cnt[slot] = ok ? (int64_t)total : 0;
        s_base = base;
        s_ok = ok;
      }
      __syncthreads();
    }
  }
}

// Last-resort finish: the range is already fully sorted in global memory (recursive partition levels + kSortOnly);
// one thread walks one range.
template <int W, int MODE>
__global__ void k_serial(LocalArgs a, int nwork) {
  const int wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= nwork) return;
  const int slot = a.work[wi].slot;
  const int64_t n = a.bkt_size[slot];
  const GmemAcc<W> acc{a.in + a.bkt_start[slot] * (int64_t)W};
  if constexpr (MODE == kCountEmit) {
    uint32_t total = 0;
    for (int64_t i = 0, j; i < n; i = j) {
      j = i + 1;
      while (j < n && cmp_rec<W>(acc(j), acc(i)) == 0) ++j;
      total += (j - i) >= a.min_count;
    }
    unsigned long long base = atomicAdd(a.arena_cursor, (unsigned long long)total);
    const bool ok = base + total <= a.arena_cap;
    if (!ok) atomicExch(a.overflow_flag, 1);
    a.desc_off[slot] = (int64_t)base;
    a.desc_cnt[slot] = ok ? (int64_t)total : 0;
    uint32_t o = 0;
    for (int64_t i = 0, j; i < n; i = j) {
      j = i + 1;
      while (j < n && cmp_rec<W>(acc(j), acc(i)) == 0) ++j;
      const uint32_t c = (j - i) > kMaxMul ? (uint32_t)kMaxMul : (uint32_t)(j - i);
      if (a.counting) atomicAdd(a.counting + c, 1ull);
      if ((j - i) >= a.min_count && ok) {
        write_edge<W>(a.arena + (base + o) * (unsigned long long)a.words_edge, acc(i), a.words_edge, c);
        ++o;
      }
    }
  } else if constexpr (MODE == kSdbgEmit) {
    SdbgTally t = sdbg_walk<W, false>(acc, (int64_t)0, n, a.k, a.words_tip, a.tip_mode, nullptr, nullptr, a.bucket_stats);
    const unsigned long long ib = atomicAdd(a.sd_cursor, (unsigned long long)t.items);
    const unsigned long long tb = atomicAdd(a.sd_cursor + 1, (unsigned long long)t.tips);
    const bool ok = ib + t.items <= a.sd_item_cap && tb + t.tips <= a.sd_tip_cap;
    if (!ok) atomicExch(a.overflow_flag, 1);
    a.sd_item_off[slot] = (int64_t)ib;
    a.sd_tip_off[slot] = (int64_t)tb;
    a.sd_items[slot] = ok ? (int64_t)t.items : 0;
    a.sd_tips[slot] = ok ? (int64_t)t.tips : 0;
    a.sd_large[slot] = ok ? (int64_t)t.large : 0;
    if (ok)
      sdbg_walk<W, true>(acc, (int64_t)0, n, a.k, a.words_tip, a.tip_mode, a.sd_rec + ib, a.sd_labels + tb * (unsigned long long)a.words_tip);
  }
}

}  // namespace mf
