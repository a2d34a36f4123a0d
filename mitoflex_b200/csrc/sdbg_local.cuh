// sdbg_local.cuh -- per-bucket BOSS emission (SeqToSdbg::Lv2Postprocess), one CTA per bucket of sdbg items.
//
// sdbg items are almost all distinct and uniformly spread over the bits below the bucket prefix, so the bucket is sorted
// by ONE counting split on the next <= 11 bits (shared atomics give an arbitrary slot inside the sub-bin) followed by a
// rank fix inside each sub-bin (an item's final slot = sub-bin start + number of smaller items in the sub-bin; sub-bins
// hold one or two items).  Every thread then walks a contiguous run of whole (k-1)-prefix groups ONCE, staging the emitted
// records in shared memory; one block scan gives the offsets for a coalesced copy-out.  Buckets with a crowded sub-bin
// (low-complexity sequence) or more items than fit go to the bail list and take the general LSD kernel (local.cuh).
#pragma once
#include "common.cuh"
#include "local.cuh"

namespace mf {

constexpr int kSdNT = 512;
constexpr int kSdSubBitsMax = 11;
constexpr int kSdSubMax = 384;   // items one sub-bin may hold on this path (error neighbours of deep-coverage k-mers cluster)

inline size_t sdbg_fixed_words() { return (size_t)(1 << kSdSubBitsMax) + 2 + (kSdNT + 2) + 48 + 16; }
inline int sdbg_cap(int W) {
  const int budget = 110 * 1024 - (int)sdbg_fixed_words() * 4;
  int cap = budget / (8 * W + 2);
  cap &= ~1;
  return cap > 65534 ? 65534 : cap;
}
inline size_t sdbg_smem_bytes(int W, int cap) { return ((size_t)cap * W * 2 + (size_t)cap / 2 + sdbg_fixed_words()) * 4; }

// items compare on everything but the 16 multiplicity bits of the last word
template <int W>
__device__ __forceinline__ bool item_less(const uint32_t *x, const uint32_t *y) {
#pragma unroll
  for (int i = 0; i < W - 1; ++i) {
    const uint32_t p = x[i], q = y[i];
    if (p != q) return p < q;
  }
  return (x[W - 1] & 0xffff0000u) < (y[W - 1] & 0xffff0000u);
}
template <int W>
__device__ __forceinline__ bool item_same(const uint32_t *x, const uint32_t *y) {
#pragma unroll
  for (int i = 0; i < W - 1; ++i)
    if (x[i] != y[i]) return false;
  return (x[W - 1] & 0xffff0000u) == (y[W - 1] & 0xffff0000u);
}

// SeqToSdbg::Lv2Postprocess over sorted items [b, e) forming whole (k-1)-prefix groups; emit(first item of the run, w, last,
// is_dollar, multiplicity, b) once per output item, in order.
template <int W, class Emit>
__device__ __forceinline__ void sdbg_walk_emit(const uint32_t *rec, int b, int e, int k, Emit &&emit) {
  for (int gs = b, ge; gs < e; gs = ge) {
    const uint32_t *g0 = rec + (size_t)gs * W;
    ge = gs + 1;
    while (ge < e && !item_diff_km1<W>(g0, rec + (size_t)ge * W, k)) ++ge;
    int has_solid_a = 0, has_solid_b = 0, outputed_b = 0;
    int last_a0 = -1, last_a1 = -1, last_a2 = -1, last_a3 = -1;
    for (int i = gs; i < ge; ++i) {
      const uint32_t *it = rec + (size_t)i * W;
      const int ca = item_a<W>(it, k), cb = item_b<W>(it);
      if (ca != kSentinel && cb != kSentinel) { has_solid_a |= 1 << ca; has_solid_b |= 1 << cb; }
      if (ca != kSentinel && (cb != kSentinel || !(has_solid_a & (1 << ca)))) {
        if (ca == 0) last_a0 = i; else if (ca == 1) last_a1 = i; else if (ca == 2) last_a2 = i; else last_a3 = i;
      }
    }
    for (int i = gs, j; i < ge; i = j) {
      const uint32_t *it = rec + (size_t)i * W;
      const int ca = item_a<W>(it, k), cb = item_b<W>(it);
      j = i + 1;
      uint32_t inv_mul = it[W - 1] & 0xffffu;   // 65535 - multiplicity: the smallest wins
      while (j < ge) {
        const uint32_t *nx = rec + (size_t)j * W;
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
      const int la = ca == 0 ? last_a0 : ca == 1 ? last_a1 : ca == 2 ? last_a2 : last_a3;
      const int last = (ca == kSentinel) ? 0 : (la == j - 1 ? 1 : 0);
      emit(i, w, last, is_dollar, kMaxMul - (int)inv_mul, cb);
    }
  }
}

template <int W>
__global__ void __launch_bounds__(kSdNT) k_sdbg_local(LocalArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NT = kSdNT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cap = a.cap;
  uint32_t *recA = smem;                                                   // [cap*W] loaded items, later the sorted items
  uint32_t *recB = recA + (size_t)cap * W;                                 // [cap*W] split order; later stage u32[cap] | src u16[cap]
  uint16_t *rk = reinterpret_cast<uint16_t *>(recB + (size_t)cap * W);     // [cap]
  uint32_t *bins = reinterpret_cast<uint32_t *>(rk + cap);                 // [2049]
  uint32_t *strip = bins + (1 << kSdSubBitsMax) + 2;                       // [NT+1]
  uint32_t *scratch = strip + NT + 2;                                      // [34]
  int *s_flag = reinterpret_cast<int *>(scratch + 48);                     // [16]

  const int slot = a.work ? a.work[blockIdx.x].slot : (int)blockIdx.x;
  const int64_t start = a.bkt_start[slot];
  const int64_t n64 = a.bkt_size[slot];
  if (n64 == 0) return;
  auto bail = [&]() {
    if (tid == 0) {
      const int p = atomicAdd(a.bail_count, 1);
      a.bail_list[p] = slot;
    }
  };
  if (n64 > cap) {
    bail();
    if (tid == 0) atomicAdd(a.overflow_flag + 1, 1);   // statistics: too many items
    return;
  }
  const int n = (int)n64;
  {
    const uint32_t *src = a.in + start * (int64_t)W;
    const int nw = n * W;
    if constexpr (W % 2 == 0) {
      const uint2 *s2 = reinterpret_cast<const uint2 *>(src);
      uint2 *d2 = reinterpret_cast<uint2 *>(recA);
      for (int i = tid; i < nw / 2; i += NT) d2[i] = s2[i];
    } else {
      for (int i = tid; i < nw; i += NT) recA[i] = src[i];
    }
  }
  int sub_bits = 4;
  while ((1 << sub_bits) < n && sub_bits < kSdSubBitsMax) ++sub_bits;
  sub_bits = min(sub_bits, a.sort_bits - a.bit_off);
  const int nsb = 1 << sub_bits;
  for (int i = tid; i <= nsb; i += NT) bins[i] = 0;
  if (tid < 16) s_flag[tid] = 0;
  __syncthreads();
  // ---- 1. counting split on the next sub_bits
  bool crowded = false;
  for (int i = tid; i < n; i += NT) {
    const uint32_t r = atomicAdd(bins + rec_digit_mem<W>(recA + (size_t)i * W, a.bit_off, sub_bits), 1u);
    rk[i] = (uint16_t)r;
    crowded |= r >= (uint32_t)kSdSubMax;
  }
  if (__syncthreads_or(crowded)) {
    bail();
    if (tid == 0) atomicAdd(a.overflow_flag + 2, 1);   // statistics: crowded sub-bin
    return;
  }
  block_excl_scan<NT>(bins, nsb + 1, scratch);
  for (int i = tid; i < n; i += NT) {
    const uint32_t *r = recA + (size_t)i * W;
    const uint32_t pos = bins[rec_digit_mem<W>(r, a.bit_off, sub_bits)] + rk[i];
#pragma unroll
    for (int c = 0; c < W; ++c) recB[(size_t)pos * W + c] = r[c];
  }
  __syncthreads();
  // ---- 2. rank fix inside each sub-bin: final slot = sub-bin start + number of items ordered before this one
  for (int p = tid; p < n; p += NT) {
    uint32_t me[W];
#pragma unroll
    for (int c = 0; c < W; ++c) me[c] = recB[(size_t)p * W + c];
    const uint32_t d = rec_digit<W>(me, a.bit_off, sub_bits);
    const int b = (int)bins[d], e = (int)bins[d + 1];
    int r = 0;
    for (int q = b; q < e; ++q) {
      if (q == p) continue;
      const uint32_t *o = recB + (size_t)q * W;
      r += (item_less<W>(o, me) || (q < p && !item_less<W>(me, o))) ? 1 : 0;
    }
#pragma unroll
    for (int c = 0; c < W; ++c) recA[(size_t)(b + r) * W + c] = me[c];
  }
  __syncthreads();
  // ---- 3. every thread walks a contiguous run of whole (k-1)-prefix groups once
  uint32_t *stage = recB;                                              // [cap] emitted records, at the run's item positions
  uint16_t *srcp = reinterpret_cast<uint16_t *>(recB + cap);          // [cap] source item of each emitted record
  const int ipt = (n + NT - 1) / NT;
  auto next_head = [&](int p) {   // first group head at or after p
    while (p > 0 && p < n && !item_diff_km1<W>(recA + (size_t)p * W, recA + (size_t)(p - 1) * W, a.k)) ++p;
    return min(p, n);
  };
  const int s = next_head(min(n, tid * ipt)), e = next_head(min(n, (tid + 1) * ipt));
  const bool one_bucket = a.bit_off >= 16;   // all items of the CTA share the megahit bucket (first 8 bases)
  uint32_t cnt = 0, tips = 0, large = 0;
  sdbg_walk_emit<W>(recA, s, e, a.k, [&](int i, int w, int last, int is_dollar, int mul, int cb) {
    stage[s + cnt] = (uint32_t)w | ((uint32_t)last << 4) | ((uint32_t)is_dollar << 5) | ((uint32_t)mul << 8);
    srcp[s + cnt] = (uint16_t)i;
    ++cnt;
    tips += is_dollar;
    large += mul > 254;
    if (!one_bucket) {
      unsigned long long *bs = a.bucket_stats + (size_t)(recA[(size_t)i * W] >> 16) * 3;
      atomicAdd(bs, 1ull);
      if (is_dollar) atomicAdd(bs + 1, 1ull);
      if (mul > 254) atomicAdd(bs + 2, 1ull);
    }
  });
  strip[tid] = cnt | (tips << 16);   // cap < 65536, so neither half can carry into the other
  if (tid == 0) strip[NT] = 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) large += __shfl_xor_sync(0xffffffffu, large, o);
  if (lane == 0) scratch[36 + warp] = large;
  __syncthreads();
  const uint32_t tot = block_excl_scan<NT>(strip, NT + 1, scratch);
  const uint32_t tot_items = tot & 0xffffu, tot_tips = tot >> 16;
  if (tid == 0) {
    uint32_t L = 0;
    for (int w2 = 0; w2 < NT / 32; ++w2) L += scratch[36 + w2];
    const unsigned long long ib = atomicAdd(a.sd_cursor, (unsigned long long)tot_items);
    const unsigned long long tb = atomicAdd(a.sd_cursor + 1, (unsigned long long)tot_tips);
    const int ok = ib + tot_items <= a.sd_item_cap && tb + tot_tips <= a.sd_tip_cap;
    if (!ok) atomicExch(a.overflow_flag, 1);
    a.sd_item_off[slot] = (int64_t)ib;
    a.sd_tip_off[slot] = (int64_t)tb;
    a.sd_items[slot] = ok ? (int64_t)tot_items : 0;
    a.sd_tips[slot] = ok ? (int64_t)tot_tips : 0;
    a.sd_large[slot] = ok ? (int64_t)L : 0;
    if (one_bucket && ok) {
      unsigned long long *bs = a.bucket_stats + (size_t)(recA[0] >> 16) * 3;
      if (tot_items) atomicAdd(bs, (unsigned long long)tot_items);
      if (tot_tips) atomicAdd(bs + 1, (unsigned long long)tot_tips);
      if (L) atomicAdd(bs + 2, (unsigned long long)L);
    }
    s_flag[1] = ok;
    s_flag[2] = (int)(uint32_t)ib; s_flag[3] = (int)(uint32_t)(ib >> 32);
    s_flag[4] = (int)(uint32_t)tb; s_flag[5] = (int)(uint32_t)(tb >> 32);
  }
  __syncthreads();
  if (!s_flag[1]) return;
  const unsigned long long item_base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
  const unsigned long long tip_base = ((unsigned long long)(uint32_t)s_flag[5] << 32) | (uint32_t)s_flag[4];
  const uint32_t my_off = strip[tid] & 0xffffu, my_tip = strip[tid] >> 16;
  // tip labels (rare) straight from the sorted items
  if (tips) {
    const int Wt = a.words_tip;
    uint32_t ot = 0;
    for (uint32_t j = 0; j < cnt; ++j) {
      if (!((stage[s + j] >> 5) & 1u)) continue;
      const uint32_t *it = recA + (size_t)srcp[s + j] * W;
      uint32_t *lab = a.sd_labels + (tip_base + my_tip + ot) * (unsigned long long)Wt;
      for (int q = 0; q < Wt; ++q) lab[q] = it[q];
      if (a.tip_mode == 1) {
        // read2sdbg stage-2 items carry only flag<<3 | b below the bases (Read2SdbgS2::Lv2ExtractSubString)
        if (W == Wt) lab[Wt - 1] &= 0xfff00000u;
        if ((2 * a.k + 4 + 31) / 32 == Wt) lab[Wt - 1] |= (uint32_t)item_b<W>(it);
      }
      ++ot;
    }
  }
  __syncthreads();   // the sorted items are dead: compact the staged records over them
  uint32_t *cstage = recA;
  for (uint32_t j = 0; j < cnt; ++j) cstage[my_off + j] = stage[s + j];
  __syncthreads();
  uint32_t *dst = a.sd_rec + item_base;
  for (uint32_t j = tid; j < tot_items; j += NT) dst[j] = cstage[j];
}

}  // namespace mf
