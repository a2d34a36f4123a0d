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
//   kSdbgCount / kSdbgEmit : BOSS emission per (k-1)-prefix group (SeqToSdbg::Lv2Postprocess), one thread per group run
// k_serial runs the same walkers single-threaded over an already sorted global range (last-resort path).
#pragma once
#include "common.cuh"

namespace mf {

enum LocalMode { kCountEmit = 0, kSortOnly = 1, kSdbgCount = 2, kSdbgEmit = 3, kCountPairs = 4, kCountMerge = 5 };

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
  const int64_t *sd_item_off;            // [nslots] (emit phase) global item offset
  const int64_t *sd_tip_off;             // [nslots]
  uint32_t *sd_rec;                      // emitted: w | last<<4 | tip<<5 | mult<<8
  uint32_t *sd_labels;                   // words_tip per tip
};

constexpr int kLocalDigitBits = 10;
constexpr int kLocalBins = 1 << kLocalDigitBits;
constexpr int kLocalNT = 256;

// shared memory: rec[cap*WR] | idxA,idxB,rk u16[cap] | aux u32[weighted ? cap+2 : 2NT+2] | whist u16[NWARP][1024] | bins u32[1025] | scratch
inline size_t local_smem_bytes(int WR, int cap, bool weighted) {
  size_t aux = weighted ? (size_t)cap + 2 : 2 * kLocalNT + 2;
  size_t words = (size_t)cap * WR + 3 * ((size_t)cap / 2) + aux + (size_t)(kLocalNT / 32) * kLocalBins / 2 + kLocalBins + 1 + 48;
  return words * 4;
}
// records per CTA such that two CTAs fit an SM (110 KB each)
inline int local_cap(int WR, bool weighted) {
  const int fixed = (kLocalNT / 32) * kLocalBins * 2 + (kLocalBins + 1) * 4 + 48 * 4 + (weighted ? 8 : (2 * kLocalNT + 2) * 4);
  int cap = (110 * 1024 - fixed) / (WR * 4 + 6 + (weighted ? 4 : 0));
  cap &= ~1;
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
      while (j < ge) {
        const uint32_t *nx = acc(j);
        if (item_a<W>(nx, k) != ca || item_b<W>(nx) != cb) break;
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
      const int mul = kMaxMul - (int)(it[W - 1] & 0xffffu);
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
template <int WR, int W, int NT>
__device__ __forceinline__ bool lsd_pass(const uint32_t *rec, const uint16_t *idx_in, uint16_t *idx_out, uint16_t *rk, int n,
                                         int bit_lo, int nbits, uint16_t *whist, uint32_t *bins, uint32_t *scratch) {
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
      const int r = idx_in ? idx_in[p] : p;
      d = rec_digit_mem<W>(rec + (size_t)r * WR, bit_lo, nbits);
    }
    const unsigned m = match_digit(d, valid);
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
      const int r = idx_in ? idx_in[p] : p;
      const uint32_t d = rec_digit_mem<W>(rec + (size_t)r * WR, bit_lo, nbits);
      idx_out[bins[d] + wh[d] + rk[p]] = (uint16_t)r;
    }
  }
  __syncthreads();
  return true;
}

// Block-wide compaction of flagged positions in [0, n): list[j] = j-th flagged position (ascending); returns the count.
// flag(p) is evaluated twice.  Threads own contiguous strips, so the list is ordered.
template <int NT, class F>
__device__ __forceinline__ uint32_t compact_positions(int n, F flag, uint16_t *list, uint32_t *strip_cnt /*[NT+1]*/,
                                                      uint32_t *scratch) {
  const int tid = threadIdx.x;
  const int per = (n + NT - 1) / NT;
  const int b = min(n, tid * per), e = min(n, b + per);
  uint32_t c = 0;
  for (int p = b; p < e; ++p) c += flag(p) ? 1u : 0u;
  strip_cnt[tid] = c;
  if (tid == 0) strip_cnt[NT] = 0;
  __syncthreads();
  const uint32_t total = block_excl_scan<NT>(strip_cnt, NT + 1, scratch);
  uint32_t o = strip_cnt[tid];
  for (int p = b; p < e; ++p)
    if (flag(p)) list[o++] = (uint16_t)p;
  __syncthreads();
  return total;
}

template <int W, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_local(LocalArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NWARP = NT / 32;
  constexpr bool WEIGHTED = MODE == kCountMerge;
  constexpr int WR = WEIGHTED ? W + 1 : W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cap = a.cap;

  uint32_t *rec = smem;                                                       // [cap*WR]
  uint16_t *idxA = reinterpret_cast<uint16_t *>(rec + (size_t)cap * WR);      // [cap]
  uint16_t *idxB = idxA + cap;                                                // [cap]
  uint16_t *rk = idxB + cap;                                                  // [cap]
  uint32_t *aux = reinterpret_cast<uint32_t *>(rk + cap);                     // [WEIGHTED ? cap+2 : 2NT+2]
  uint16_t *whist = reinterpret_cast<uint16_t *>(aux + (WEIGHTED ? cap + 2 : 2 * NT + 2));   // [NWARP][1024]
  uint32_t *bins = reinterpret_cast<uint32_t *>(whist + NWARP * kLocalBins);  // [1025]
  uint32_t *scratch = bins + kLocalBins + 1;                                  // [34]
  int *s_flag = reinterpret_cast<int *>(scratch + 34);                        // [8]

  // ---- work item
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
  // ---- 1. load
  if constexpr (MODE == kCountMerge) {
    // `start` / `n64` index pieces: gather every chunk's pairs
    const int first = (int)start, np = (int)n64;
    int total = 0;
    for (int q = 0; q < np; ++q) total += a.piece_cnt[first + q];
    if (total > cap) {
      if (tid == 0) {
        int p = atomicAdd(a.bail_count, 1);
        a.bail_list[p] = slot;
      }
      return;
    }
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
    if (n64 > cap) {                       // does not fit: the chunked / fallback path takes it
      if (tid == 0) {
        int p = atomicAdd(a.bail_count, 1);
        a.bail_list[p] = slot;
      }
      return;
    }
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
  if (tid < 8) s_flag[tid] = 0;
  __syncthreads();

  // ---- 2. LSD sort of the index array over bits [bit_off, sort_bits), least significant digit first
  const uint16_t *cur = nullptr;   // nullptr == identity order
  {
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
  uint16_t *list = (cur == idxA) ? idxB : idxA;   // the free index array: head / group positions
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

  if constexpr (MODE == kCountEmit || MODE == kCountPairs || MODE == kCountMerge) {
    // ---- 3. runs of equal keys
    auto is_head = [&](int p) { return p == 0 || cmp_rec<W>(acc(p), acc(p - 1)) != 0; };
    const uint32_t nh = compact_positions<NT>(n, is_head, list, strip, scratch);
    if constexpr (WEIGHTED) {
      // aux[p] = sum of the counts of sorted positions < p (pairs carry their chunk's count in word W)
      const int per = (n + NT - 1) / NT;
      const int b = min(n, tid * per), e = min(n, b + per);
      uint32_t s = 0;
      for (int p = b; p < e; ++p) s += acc(p)[W];
      strip[tid] = s;
      if (tid == 0) strip[NT] = 0;
      __syncthreads();
      const uint32_t tot = block_excl_scan<NT>(strip, NT + 1, scratch);
      uint32_t run = strip[tid];
      for (int p = b; p < e; ++p) { aux[p] = run; run += acc(p)[W]; }
      if (tid == 0) aux[n] = tot;
      __syncthreads();
    }
    auto run_count = [&](uint32_t j) -> uint32_t {
      const int p0 = list[j], p1 = (j + 1 < nh) ? (int)list[j + 1] : n;
      if constexpr (WEIGHTED) return aux[p1] - aux[p0];
      else return (uint32_t)(p1 - p0);
    };
    if constexpr (MODE == kCountPairs) {
      // every distinct key of this chunk with its count -> pair arena
      if (tid == 0) {
        unsigned long long base = atomicAdd(a.pair_cursor, (unsigned long long)nh);
        int ok = base + nh <= a.pair_cap;
        if (!ok) atomicExch(a.overflow_flag + 1, 1);
        a.piece_off[blockIdx.x] = (int64_t)base;
        a.piece_cnt[blockIdx.x] = ok ? (int32_t)nh : 0;
        s_flag[1] = ok;
        s_flag[2] = (int)(uint32_t)base;
        s_flag[3] = (int)(uint32_t)(base >> 32);
      }
      __syncthreads();
      if (!s_flag[1]) return;
      const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
      for (uint32_t j = tid; j < nh; j += NT) {
        const uint32_t *key = acc(list[j]);
        uint32_t *dst = a.pair_arena + (base + j) * (unsigned long long)(W + 1);
#pragma unroll
        for (int t = 0; t < W; ++t) dst[t] = key[t];
        dst[W] = run_count(j);
      }
      return;
    } else {
      // ---- 4. solid filter + emission; rk is free after the sort and takes the solid-run list
      uint16_t *solid = rk;
      auto is_solid = [&](int j) { return run_count((uint32_t)j) >= (uint32_t)a.min_count; };
      const uint32_t ns = compact_positions<NT>((int)nh, is_solid, solid, strip, scratch);
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
      __syncthreads();
      if (a.counting) {
        // distinct-edge multiplicity histogram (<prefix>.counting): warp-aggregate equal counts before the atomic
        for (uint32_t j = tid; j < ((nh + 31) & ~31u); j += NT) {
          const bool v = j < nh;
          const uint32_t c = v ? min(run_count(j), (uint32_t)kMaxMul) : 0u;
          const unsigned m = match_digit(c, v);
          if (v && lane == (unsigned)(__ffs(m) - 1)) atomicAdd(a.counting + c, (unsigned long long)__popc(m));
        }
      }
      if (!s_flag[1]) return;
      const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
      const int We = a.words_edge;
      for (uint32_t q = tid; q < ns; q += NT) {
        const uint32_t j = solid[q];
        write_edge<W>(a.arena + (base + q) * (unsigned long long)We, acc(list[j]), We, run_count(j));
      }
      return;
    }
  }

  if constexpr (MODE == kSdbgCount || MODE == kSdbgEmit) {
    // ---- groups of equal (k-1)-prefix; threads own contiguous runs of groups (groups are a handful of items)
    auto is_ghead = [&](int p) { return p == 0 || item_diff_km1<W>(acc(p), acc(p - 1), a.k); };
    const uint32_t ng = compact_positions<NT>(n, is_ghead, list, strip, scratch);
    uint32_t my_items = 0, my_tips = 0, my_large = 0;
    const int gper = ((int)ng + NT - 1) / NT;
    const int gb = min((int)ng, tid * gper), ge = min((int)ng, gb + gper);
    for (int g = gb; g < ge; ++g) {
      const int p0 = list[g], p1 = (g + 1 < (int)ng) ? (int)list[g + 1] : n;
      SdbgTally t = sdbg_walk<W, false>(acc, p0, p1, a.k, a.words_tip, a.tip_mode, nullptr, nullptr,
                                        MODE == kSdbgCount ? a.bucket_stats : nullptr);
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
    if constexpr (MODE == kSdbgCount) {
#pragma unroll
      for (int o = 16; o; o >>= 1) my_large += __shfl_xor_sync(0xffffffffu, my_large, o);
      if (lane == 0) scratch[warp] = my_large;
      __syncthreads();
      if (tid == 0) {
        uint32_t L = 0;
        for (int w = 0; w < NWARP; ++w) L += scratch[w];
        a.sd_items[slot] = (int64_t)tot_items;
        a.sd_tips[slot] = (int64_t)tot_tips;
        a.sd_large[slot] = (int64_t)L;
      }
    } else {
      const int64_t item_base = a.sd_item_off[slot] + s_items[tid], tip_base = a.sd_tip_off[slot] + s_tips[tid];
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
  } else if constexpr (MODE == kSdbgCount) {
    SdbgTally t = sdbg_walk<W, false>(acc, (int64_t)0, n, a.k, a.words_tip, a.tip_mode, nullptr, nullptr, a.bucket_stats);
    a.sd_items[slot] = (int64_t)t.items;
    a.sd_tips[slot] = (int64_t)t.tips;
    a.sd_large[slot] = (int64_t)t.large;
  } else if constexpr (MODE == kSdbgEmit) {
    sdbg_walk<W, true>(acc, (int64_t)0, n, a.k, a.words_tip, a.tip_mode, a.sd_rec + a.sd_item_off[slot],
                       a.sd_labels + a.sd_tip_off[slot] * (int64_t)a.words_tip);
  }
}

}  // namespace mf
