// items.cuh -- sdbg item generation (SeqToSdbg::Lv0CalcBucketSize / Lv2ExtractSubString).
// Every stored sequence of length L >= k+1 yields, on both strands, items at offsets 0..L-k+1:
//   "$xxxx, xxxxx, ..., xxxx$"  = k bases (k-1 for the last), the preceding base b (or $), the multiplicity
// (only for real edges).  Item word layout: bases left aligned, last word low 20 bits =
// flag(k bases present)<<19 | b<<16 | (65535 - multiplicity), so a plain ascending sort over all words puts
// larger multiplicities first.
#pragma once
#include "common.cuh"

namespace mf {

// shift a WS-word left-aligned base string left by `chars` bases, keep `nchars` bases, emit WD words
template <int WS, int WD>
__device__ __forceinline__ void window_item(const uint32_t (&t)[WS], int chars, int nchars, uint32_t flag, uint32_t prev,
                                            uint32_t cnt, uint32_t (&out)[WD]) {
  const int sh = 2 * chars;   // 0, 2 or 4
#pragma unroll
  for (int i = 0; i < WD; ++i) {
    uint32_t hi = i < WS ? t[i] : 0u, lo = (i + 1) < WS ? t[i + 1] : 0u;
    out[i] = __funnelshift_l(lo, hi, sh);
  }
  const int nb = 2 * nchars, wm = nb >> 5, rem = nb & 31;
#pragma unroll
  for (int i = 0; i < WD; ++i) {
    if (i > wm) out[i] = 0;
    else if (i == wm) out[i] &= rem ? (0xffffffffu << (32 - rem)) : 0u;
  }
  out[WD - 1] |= (flag << 19) | (prev << 16) | (uint32_t)(kMaxMul - (int)cnt);
}

// reverse complement of `nchars` left-aligned bases held in WS words
template <int WS>
__device__ __forceinline__ void revcomp_words(const uint32_t (&t)[WS], int nchars, uint32_t (&out)[WS]) {
  const int nb = 2 * nchars, wm = nb >> 5, rem = nb & 31;
  const int pad = 32 * WS - nb, pw = pad >> 5, pb = pad & 31;
  uint32_t rr[WS];
#pragma unroll
  for (int i = 0; i < WS; ++i) {
    const int s = WS - 1 - i;
    uint32_t c = ~t[s];
    if (s > wm) c = 0;                                         // word entirely inside the pad
    else if (s == wm) c &= rem ? (0xffffffffu << (32 - rem)) : 0u;
    uint32_t x = __brev(c);
    rr[i] = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
  }
  // the reversed string sits right aligned: shift left by `pad` bits across words
#pragma unroll
  for (int i = 0; i < WS; ++i) {
    uint32_t hi = 0, lo = 0;
#pragma unroll
    for (int q = 0; q < WS; ++q) {
      if (q == i + pw) hi = rr[q];
      if (q == i + pw + 1) lo = rr[q];
    }
    out[i] = __funnelshift_l(lo, hi, pb);
  }
}

// One thread per edge record ((k+1)-mer + multiplicity in the low 16 bits of the last word): 6 items.
// WK = words of the (k+1)-mer, WE = words of the edge record, WI = words of an item.
template <int WK, int WE, int WI>
__global__ void k_items_from_edges(const uint32_t *__restrict__ edges, int64_t n_edges, int k, uint32_t *__restrict__ items) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  uint32_t fw[WK], rc[WK];
  const uint32_t *src = edges + e * WE;
#pragma unroll
  for (int i = 0; i < WK; ++i) fw[i] = src[i];
  const uint32_t mult = src[WE - 1] & 0xffffu;
  const int pad = 32 * WK - 2 * (k + 1);
  fw[WK - 1] &= 0xffffffffu << pad;        // drop the multiplicity if it shares the last key word
  revcomp_words<WK>(fw, k + 1, rc);
  uint32_t *dst = items + e * 6 * WI;
#pragma unroll
  for (int strand = 0; strand < 2; ++strand) {
    const uint32_t(&t)[WK] = strand ? rc : fw;
    const uint32_t c0 = t[0] >> 30, c1 = (t[0] >> 28) & 3;
    uint32_t it[WI];
    window_item<WK, WI>(t, 0, k, 1, kSentinel, 0, it);
#pragma unroll
    for (int i = 0; i < WI; ++i) dst[i] = it[i];
    window_item<WK, WI>(t, 1, k, 1, c0, mult, it);
#pragma unroll
    for (int i = 0; i < WI; ++i) dst[WI + i] = it[i];
    window_item<WK, WI>(t, 2, k - 1, 0, c1, 0, it);
#pragma unroll
    for (int i = 0; i < WI; ++i) dst[2 * WI + i] = it[i];
    dst += 3 * WI;
  }
}


// ---------------------------------------------------------------------------------------------------------
// Helpers of the item filter (kmerset.cuh): k-mers of up to 64 / 128 bits, left aligned.  Of the 6 items of an edge only the 2
// real ones always reach the graph; a "$"-head item survives Lv2Postprocess only if its k-mer has no incoming solid edge, a
// "$"-tail item only if its k-mer has no outgoing one -- rare at assembly depths.  Items that the walker would drop anyway
// are never generated, sorted or walked.  Dropping them cannot change any other output: they carry no multiplicity, never
// feed has_solid_a/b or last_a, and whole (a, b) runs of them are skipped together.
__device__ __forceinline__ unsigned long long revcomp64(unsigned long long t, int nchars) {   // left-aligned 2*nchars bits
  unsigned long long x = __brevll(~t);   // complement, reversed bit order: bases reversed with their two bits swapped
  x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
  return x << (64 - 2 * nchars);         // the reversed string sat right aligned
}
// ---- 32 <= k <= 63: the k-mer takes up to 126 bits, the table holds 16-byte slots (ATOMG.CAS.128 on sm_100a)
struct K128 {
  unsigned long long hi, lo;
};
__device__ __forceinline__ K128 shl128(K128 a, int s) {   // 0 <= s < 128
  if (s == 0) return a;
  if (s >= 64) return K128{a.lo << (s - 64), 0ull};
  return K128{(a.hi << s) | (a.lo >> (64 - s)), a.lo << s};
}
__device__ __forceinline__ K128 mask_top128(K128 a, int bits) {   // keep the top `bits` bits, 0 < bits <= 128
  if (bits >= 128) return a;
  if (bits >= 64) return K128{a.hi, bits == 64 ? 0ull : (a.lo & (~0ull << (128 - bits)))};
  return K128{a.hi & (~0ull << (64 - bits)), 0ull};
}
__device__ __forceinline__ unsigned long long rev_bases64(unsigned long long x) {
  x = __brevll(x);
  return ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
}
__device__ __forceinline__ K128 revcomp128(K128 t, int nchars) {   // left-aligned 2*nchars bits
  const K128 r{rev_bases64(~t.lo), rev_bases64(~t.hi)};             // whole register reversed: the string sits right aligned
  return shl128(r, 128 - 2 * nchars);
}
__device__ __forceinline__ unsigned __int128 pack128(K128 a) { return ((unsigned __int128)a.hi << 64) | a.lo; }

// General sequences (contigs etc.), stored orientation, 2-bit packed back to back.
// One thread per item; item -> sequence by binary search over item_base (sequences are long, few).
template <int WI>
__device__ __forceinline__ void seq_item(const uint32_t *__restrict__ packed, const int64_t *__restrict__ seq_start,
                                         const uint16_t *__restrict__ seq_mult, const int64_t *__restrict__ item_base, int nseq,
                                         int64_t x, int k, uint32_t (&w)[WI]) {
  int lo = 0, hi = nseq;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (item_base[mid] <= x) lo = mid; else hi = mid;
  }
  const int64_t s0 = seq_start[lo];
  const int L = (int)(seq_start[lo + 1] - s0);
  const int per = L - k + 2;
  int64_t r = x - item_base[lo];
  const int strand = r >= per;
  const int of = (int)(strand ? r - per : r);
  const int nchars = (of + k > L) ? k - 1 : k;
  const uint32_t cnt = (of > 0 && of + k <= L) ? seq_mult[lo] : 0u;
  auto base_at = [&](int64_t g) -> uint32_t { return (packed[g >> 4] >> (30 - 2 * (int)(g & 15))) & 3u; };
  // forward window start (in the stored string) covering the item's bases
  const int f = strand ? (L - of - nchars) : of;
  uint32_t raw[WI + 1];
  {
    const int64_t g = s0 + f;
    const int64_t wi = g >> 4;
    const int sh = (int)(g & 15) * 2;
#pragma unroll
    for (int i = 0; i < WI; ++i) raw[i] = __funnelshift_l(packed[wi + i + 1], packed[wi + i], sh);
  }
  const int nb = 2 * nchars, wm = nb >> 5, rem = nb & 31;
#pragma unroll
  for (int i = 0; i < WI; ++i) {
    w[i] = raw[i];
    if (i > wm) w[i] = 0;
    else if (i == wm) w[i] &= rem ? (0xffffffffu << (32 - rem)) : 0u;
  }
  uint32_t prev;
  if (strand) {
    uint32_t t[WI];
    revcomp_words<WI>(w, nchars, t);
#pragma unroll
    for (int i = 0; i < WI; ++i) w[i] = t[i];
    prev = of == 0 ? (uint32_t)kSentinel : 3u - base_at(s0 + L - of);
  } else {
    prev = of == 0 ? (uint32_t)kSentinel : base_at(s0 + of - 1);
  }
  w[WI - 1] |= ((uint32_t)(nchars == k) << 19) | (prev << 16) | (uint32_t)(kMaxMul - (int)cnt);
}
template <int WI>
__global__ void k_items_from_seqs(const uint32_t *__restrict__ packed, const int64_t *__restrict__ seq_start,
                                  const uint16_t *__restrict__ seq_mult, const int64_t *__restrict__ item_base, int nseq,
                                  int64_t n_items, int k, uint32_t *__restrict__ items) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n_items) return;
  uint32_t w[WI];
  seq_item<WI>(packed, seq_start, seq_mult, item_base, nseq, x, k, w);
  uint32_t *dst = items + x * WI;
#pragma unroll
  for (int i = 0; i < WI; ++i) dst[i] = w[i];
}

// ---------------------------------------------------------------------------------------------------------
// Ranged generation for the memory-bounded sdbg rounds (megahit bounds SeqToSdbg by --host_mem the same way: lv1 passes
// over bucket ranges).  An item belongs to a round if its top `bin_bits` bits fall into [lo, hi).  HIST = true: nothing is
// written, the per-bin item counts of the whole input go to hist[1 << bin_bits] (shared histogram per CTA, flushed once);
// HIST = false: items of the range are appended through a warp-aggregated cursor (arbitrary order, like the filtered path).
constexpr int kRangedNT = 256;
template <int WI, int N>
__device__ __forceinline__ void ranged_sink(const uint32_t (&it)[N][WI], const bool (&in)[N], uint32_t *__restrict__ items,
                                            unsigned long long *cursor) {
  int mine = 0;
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (in[j]) ++mine;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += v;
  }
  const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if ((threadIdx.x & 31) == 31 && warp_total) base = atomicAdd(cursor, (unsigned long long)warp_total);
  base = __shfl_sync(0xffffffffu, base, 31);
  uint32_t *dst = items + (base + (unsigned long long)(incl - mine)) * WI;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    if (in[j]) {
#pragma unroll
      for (int i = 0; i < WI; ++i) dst[i] = it[j][i];
      dst += WI;
    }
  }
}
template <int WK, int WE, int WI, bool HIST>
__global__ void __launch_bounds__(kRangedNT) k_items_from_edges_ranged(const uint32_t *__restrict__ edges, int64_t n_edges, int k,
                                                                        int bin_bits, uint32_t lo, uint32_t hi,
                                                                        uint32_t *__restrict__ items, unsigned long long *cursor,
                                                                        unsigned long long *__restrict__ hist) {
  extern __shared__ uint32_t sh_hist[];
  const int nbins = 1 << bin_bits;
  if constexpr (HIST) {
    for (int i = threadIdx.x; i < nbins; i += kRangedNT) sh_hist[i] = 0;
    __syncthreads();
  }
  for (int64_t base = (int64_t)blockIdx.x * kRangedNT; base < n_edges; base += (int64_t)gridDim.x * kRangedNT) {
    const int64_t e = base + threadIdx.x;
    const bool live = e < n_edges;
    uint32_t it[6][WI];
    bool in[6] = {false, false, false, false, false, false};
    if (live) {
      uint32_t fw[WK], rc[WK];
      const uint32_t *src = edges + e * WE;
#pragma unroll
      for (int i = 0; i < WK; ++i) fw[i] = src[i];
      const uint32_t mult = src[WE - 1] & 0xffffu;
      const int pad = 32 * WK - 2 * (k + 1);
      fw[WK - 1] &= 0xffffffffu << pad;        // drop the multiplicity if it shares the last key word
      revcomp_words<WK>(fw, k + 1, rc);
#pragma unroll
      for (int strand = 0; strand < 2; ++strand) {
        const uint32_t(&t)[WK] = strand ? rc : fw;
        const uint32_t c0 = t[0] >> 30, c1 = (t[0] >> 28) & 3;
        window_item<WK, WI>(t, 0, k, 1, kSentinel, 0, it[3 * strand]);
        window_item<WK, WI>(t, 1, k, 1, c0, mult, it[3 * strand + 1]);
        window_item<WK, WI>(t, 2, k - 1, 0, c1, 0, it[3 * strand + 2]);
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint32_t bin = it[j][0] >> (32 - bin_bits);
        if constexpr (HIST) atomicAdd(&sh_hist[bin], 1u);
        else in[j] = bin >= lo && bin < hi;
      }
    }
    if constexpr (!HIST) ranged_sink<WI, 6>(it, in, items, cursor);
  }
  if constexpr (HIST) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += kRangedNT)
      if (sh_hist[i]) atomicAdd(hist + i, (unsigned long long)sh_hist[i]);
  }
}
template <int WI, bool HIST>
__global__ void __launch_bounds__(kRangedNT) k_items_from_seqs_ranged(const uint32_t *__restrict__ packed,
                                                                       const int64_t *__restrict__ seq_start,
                                                                       const uint16_t *__restrict__ seq_mult,
                                                                       const int64_t *__restrict__ item_base, int nseq, int64_t n_items,
                                                                       int k, int bin_bits, uint32_t lo, uint32_t hi,
                                                                       uint32_t *__restrict__ items, unsigned long long *cursor,
                                                                       unsigned long long *__restrict__ hist) {
  extern __shared__ uint32_t sh_hist[];
  const int nbins = 1 << bin_bits;
  if constexpr (HIST) {
    for (int i = threadIdx.x; i < nbins; i += kRangedNT) sh_hist[i] = 0;
    __syncthreads();
  }
  for (int64_t base = (int64_t)blockIdx.x * kRangedNT; base < n_items; base += (int64_t)gridDim.x * kRangedNT) {
    const int64_t x = base + threadIdx.x;
    uint32_t it[1][WI];
    bool in[1] = {false};
    if (x < n_items) {
      seq_item<WI>(packed, seq_start, seq_mult, item_base, nseq, x, k, it[0]);
      const uint32_t bin = it[0][0] >> (32 - bin_bits);
      if constexpr (HIST) atomicAdd(&sh_hist[bin], 1u);
      else in[0] = bin >= lo && bin < hi;
    }
    if constexpr (!HIST) ranged_sink<WI, 1>(it, in, items, cursor);
  }
  if constexpr (HIST) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += kRangedNT)
      if (sh_hist[i]) atomicAdd(hist + i, (unsigned long long)sh_hist[i]);
  }
}

}  // namespace mf
