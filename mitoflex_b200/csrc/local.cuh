// local.cuh -- per-bucket finish: one CTA owns one bucket that fits in shared memory.
//   1. load the bucket (coalesced), 2. split it once more in shared memory on the next `sb_bits` key bits
//   (warp-private match ranking, index array only -- records stay put), 3. each thread finishes whole
//   sub-bins serially (tiny insertion sort + run-length walk), 4. block scans give output offsets.
// Consumers:
//   kCountEmit : runs of equal keys -> multiplicity, --min-count filter, edge records (KmerCounter::Lv2Postprocess)
//   kSortOnly  : write the bucket back fully sorted (used by the oversized-bucket fallback)
//   kSdbgCount / kSdbgEmit : BOSS emission per (k-1)-prefix group (SeqToSdbg::Lv2Postprocess), two phases
// The same walkers run single-threaded over an already sorted global range in k_serial (fallback path).
#pragma once
#include "common.cuh"

namespace mf {

enum LocalMode { kCountEmit = 0, kSortOnly = 1, kSdbgCount = 2, kSdbgEmit = 3 };

struct LocalArgs {
  const uint32_t *in;        // records
  uint32_t *out;             // kSortOnly: sorted records (a different buffer, same offsets)
  const int64_t *bkt_start;  // [nslots]
  const int64_t *bkt_size;   // [nslots]
  const int32_t *work;       // [nwork] slot ids handled by this launch
  int bit_off;               // bits every record of a bucket shares
  int sb_bits;               // in-bucket split width (0..kMaxDigitBits)
  int cap;                   // max records per bucket (shared memory capacity), even
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
  // --- fallback bookkeeping ---
  int32_t *bail_list;                    // slots the serial finish gave up on
  int *bail_count;
  int *overflow_flag;
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

template <int W>
inline size_t local_smem_bytes(int NT, int cap, int sb_bits) {
  size_t nsb = (size_t)1 << sb_bits;
  size_t words = (size_t)cap * W + (size_t)cap /*idx + rk, u16 each*/ + ((NT / 32) * nsb + 1) / 2 + 2 * (nsb + 1) + 40;
  return words * 4;
}

constexpr int kSerialShiftBudget = 4096;

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

// accessors over the sorted order of a range
template <int W>
struct SmemAcc {
  const uint32_t *rec;
  const uint16_t *idx;
  __device__ __forceinline__ const uint32_t *operator()(int i) const { return rec + (size_t)idx[i] * W; }
};
template <int W>
struct GmemAcc {
  const uint32_t *base;
  __device__ __forceinline__ const uint32_t *operator()(int64_t i) const { return base + i * W; }
};

// number of runs of equal records in [b, e) whose length reaches min_count
template <int W, class Acc, class I>
__device__ __forceinline__ uint32_t count_solid_runs(const Acc &acc, I b, I e, int min_count) {
  uint32_t ns = 0;
  for (I i = b, j; i < e; i = j) {
    const uint32_t *ri = acc(i);
    j = i + 1;
    while (j < e && cmp_rec<W>(acc(j), ri) == 0) ++j;
    ns += (j - i) >= (I)min_count;
  }
  return ns;
}
// KmerCounter::PackEdge for every solid run; returns records written
template <int W, class Acc, class I>
__device__ __forceinline__ uint32_t emit_solid_runs(const Acc &acc, I b, I e, int min_count, int We, uint32_t *dst,
                                                    unsigned long long *counting, bool write) {
  uint32_t o = 0;
  for (I i = b, j; i < e; i = j) {
    const uint32_t *ri = acc(i);
    j = i + 1;
    while (j < e && cmp_rec<W>(acc(j), ri) == 0) ++j;
    const int c = (j - i) > (I)kMaxMul ? kMaxMul : (int)(j - i);
    if (counting) atomicAdd(counting + c, 1ull);
    if ((j - i) >= (I)min_count && write) {
      uint32_t *d = dst + (size_t)o * We;
      for (int t = 0; t < We; ++t) d[t] = t < W ? ri[t] : 0u;
      d[We - 1] |= (uint32_t)c;
      ++o;
    }
  }
  return o;
}

struct SdbgTally {
  uint32_t items, tips, large;
};
// SeqToSdbg::Lv2Postprocess over sorted items [b, e); WRITING=false tallies, true writes at rec_out / lab_out.
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

template <int W, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_local(LocalArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NWARP = NT / 32;
  static_assert(NWARP % 2 == 0, "whist must stay word aligned");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsb = 1 << a.sb_bits;
  const int slot = a.work ? a.work[blockIdx.x] : (int)blockIdx.x;
  const int64_t start = a.bkt_start[slot];
  const int64_t n64 = a.bkt_size[slot];
  if (n64 == 0) return;
  if (n64 > a.cap) {                       // does not fit: leave it to the fallback path
    if (threadIdx.x == 0) {
      int p = atomicAdd(a.bail_count, 1);
      a.bail_list[p] = slot;
    }
    return;
  }
  const int n = (int)n64;

  uint32_t *rec = smem;                                                    // [cap*W]
  uint16_t *idx = reinterpret_cast<uint16_t *>(rec + (size_t)a.cap * W);   // [cap]
  uint16_t *rk = idx + a.cap;                                              // [cap]
  uint16_t *whist = rk + a.cap;                                            // [NWARP][nsb]
  uint32_t *sub_start = reinterpret_cast<uint32_t *>(whist + NWARP * nsb); // [nsb+1]
  uint32_t *sub_cnt = sub_start + nsb + 1;                                 // [nsb+1]
  uint32_t *scratch = sub_cnt + nsb + 1;                                   // [34]
  int *s_flag = reinterpret_cast<int *>(scratch + 34);                     // [0]=bail [1..3]=broadcast

  // 1. load
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
  for (int i = tid; i < NWARP * nsb; i += NT) whist[i] = 0;
  if (tid < 4) s_flag[tid] = 0;
  __syncthreads();

  // 2. rank on the sub-digit
  uint16_t *wh = whist + warp * nsb;
  for (int base = 0; base < n; base += NT) {
    const int i = base + tid;
    const bool valid = i < n;
    const uint32_t d = (valid && a.sb_bits) ? rec_digit_mem<W>(rec + (size_t)i * W, a.bit_off, a.sb_bits) : 0u;
    const unsigned m = match_digit(d, valid);
    const unsigned leader = (unsigned)(__ffs(m) - 1);
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = wh[d];
      wh[d] = (uint16_t)(old + __popc(m));
    }
    old = __shfl_sync(0xffffffffu, old, valid ? leader : 0);
    if (valid) rk[i] = (uint16_t)(old + __popc(m & lanemask_lt()));
    __syncwarp();
  }
  __syncthreads();
  for (int b = tid; b < nsb; b += NT) {
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
      uint32_t c = whist[w * nsb + b];
      whist[w * nsb + b] = (uint16_t)acc;
      acc += c;
    }
    sub_start[b] = acc;
  }
  if (tid == 0) sub_start[nsb] = 0;
  __syncthreads();
  block_excl_scan<NT>(sub_start, nsb + 1, scratch);   // sub_start[nsb] = n
  for (int base = 0; base < n; base += NT) {
    const int i = base + tid;
    if (i < n) {
      const uint32_t d = a.sb_bits ? rec_digit_mem<W>(rec + (size_t)i * W, a.bit_off, a.sb_bits) : 0u;
      idx[sub_start[d] + wh[d] + rk[i]] = (uint16_t)i;   // record i was ranked by this same thread/warp
    }
  }
  __syncthreads();

  // 3. serial finish of whole sub-bins: insertion sort of the index range
  for (int s = tid; s < nsb; s += NT) {
    const int b = (int)sub_start[s], e = (int)sub_start[s + 1];
    int budget = kSerialShiftBudget;
    for (int i = b + 1; i < e; ++i) {
      const uint16_t x = idx[i];
      const uint32_t *rx = rec + (size_t)x * W;
      int j = i;
      while (j > b && cmp_rec<W>(rec + (size_t)idx[j - 1] * W, rx) > 0) {
        idx[j] = idx[j - 1];
        --j;
        --budget;
      }
      idx[j] = x;
      if (budget < 0) { s_flag[0] = 1; break; }
    }
  }
  __syncthreads();
  if (s_flag[0]) {
    if (tid == 0) {
      int p = atomicAdd(a.bail_count, 1);
      a.bail_list[p] = slot;
    }
    return;
  }
  const SmemAcc<W> acc{rec, idx};

  if constexpr (MODE == kSortOnly) {
    uint32_t *dst = a.out + start * (int64_t)W;
    const int nw = n * W;
    for (int x = tid; x < nw; x += NT) {
      int p = x / W, c = x - p * W;
      dst[x] = rec[(size_t)idx[p] * W + c];
    }
    return;
  }

  if constexpr (MODE == kCountEmit) {
    for (int s = tid; s < nsb; s += NT)
      sub_cnt[s] = count_solid_runs<W>(acc, (int)sub_start[s], (int)sub_start[s + 1], a.min_count);
    if (tid == 0) sub_cnt[nsb] = 0;
    __syncthreads();
    const uint32_t total = block_excl_scan<NT>(sub_cnt, nsb + 1, scratch);
    if (tid == 0) {
      unsigned long long base = atomicAdd(a.arena_cursor, (unsigned long long)total);
      int ok = base + total <= a.arena_cap;
      if (!ok) atomicExch(a.overflow_flag, 1);
      a.desc_off[slot] = (int64_t)base;
      a.desc_cnt[slot] = ok ? (int64_t)total : 0;
      s_flag[1] = ok;
      s_flag[2] = (int)(uint32_t)base;
      s_flag[3] = (int)(uint32_t)(base >> 32);
    }
    __syncthreads();
    const bool ok = s_flag[1] != 0;
    const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
    for (int s = tid; s < nsb; s += NT)
      emit_solid_runs<W>(acc, (int)sub_start[s], (int)sub_start[s + 1], a.min_count, a.words_edge,
                         a.arena + (base + sub_cnt[s]) * (unsigned long long)a.words_edge, a.counting, ok);
    return;
  }

  if constexpr (MODE == kSdbgCount || MODE == kSdbgEmit) {
    // a (k-1)-prefix group never straddles a sub-bin because bit_off + sb_bits <= 2(k-1)
    uint32_t *sub_tip = reinterpret_cast<uint32_t *>(whist);   // whist is dead: reuse for per-sub-bin tip counts
    uint32_t my_large = 0;
    for (int s = tid; s < nsb; s += NT) {
      SdbgTally t = sdbg_walk<W, false>(acc, (int)sub_start[s], (int)sub_start[s + 1], a.k, a.words_tip, a.tip_mode,
                                        nullptr, nullptr, MODE == kSdbgCount ? a.bucket_stats : nullptr);
      sub_cnt[s] = t.items;
      sub_tip[s] = t.tips;
      my_large += t.large;
    }
    if (tid == 0) { sub_cnt[nsb] = 0; sub_tip[nsb] = 0; }
    __syncthreads();
    const uint32_t tot_items = block_excl_scan<NT>(sub_cnt, nsb + 1, scratch);
    const uint32_t tot_tips = block_excl_scan<NT>(sub_tip, nsb + 1, scratch);
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
      const int64_t item_base = a.sd_item_off[slot], tip_base = a.sd_tip_off[slot];
      for (int s = tid; s < nsb; s += NT)
        sdbg_walk<W, true>(acc, (int)sub_start[s], (int)sub_start[s + 1], a.k, a.words_tip, a.tip_mode,
                           a.sd_rec + item_base + sub_cnt[s], a.sd_labels + (tip_base + sub_tip[s]) * (int64_t)a.words_tip);
    }
  }
}

// Fallback finish for buckets that never fit in shared memory: the range is already fully sorted in global
// memory (by recursive partition levels + kSortOnly); one thread walks one range.
template <int W, int MODE>
__global__ void k_serial(LocalArgs a, int nwork) {
  const int wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= nwork) return;
  const int slot = a.work[wi];
  const int64_t n = a.bkt_size[slot];
  const GmemAcc<W> acc{a.in + a.bkt_start[slot] * (int64_t)W};
  if constexpr (MODE == kCountEmit) {
    const uint32_t total = count_solid_runs<W>(acc, (int64_t)0, n, a.min_count);
    unsigned long long base = atomicAdd(a.arena_cursor, (unsigned long long)total);
    const bool ok = base + total <= a.arena_cap;
    if (!ok) atomicExch(a.overflow_flag, 1);
    a.desc_off[slot] = (int64_t)base;
    a.desc_cnt[slot] = ok ? (int64_t)total : 0;
    emit_solid_runs<W>(acc, (int64_t)0, n, a.min_count, a.words_edge, a.arena + base * (unsigned long long)a.words_edge,
                       a.counting, ok);
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
