// kmerset_w.cuh -- the sdbg item filter (kmerset.cuh) for k >= 64: k-mers of 4..10 words.
//
// Same passes (histogram by hash slice, staged scatter of the 4 k-mers of every edge, in-order insert walk, in-order query walk,
// miss list), but a k-mer no longer fits a CAS: the table holds 64-bit INDICES into the scattered insert records, a slot is
// claimed with one CAS on the index, and keys are compared by reading the record the slot points at (the records of a slice
// were written moments ago and sit next to the slice's table lines in L2).  Slices are smaller (2^19 slots = 4 MB of table)
// because the records of a slice are 3-5x the size of its table lines.
#pragma once
#include "kmerset.cuh"

namespace mf {

constexpr int kKswSliceLog = 19;
constexpr unsigned long long kKswEmpty = ~0ull;

template <int WR>
__device__ __forceinline__ unsigned long long ksw_hash(const uint32_t (&t)[WR]) {
  unsigned long long x = 0x9e3779b97f4a7c15ull;
#pragma unroll
  for (int i = 0; i < WR; i += 2) {
    const unsigned long long w = ((unsigned long long)t[i] << 32) | (i + 1 < WR ? t[i + 1] : 0u);
    x = (x ^ w) * 0xbf58476d1ce4e5b9ull;
    x ^= x >> 29;
  }
  return x * 0x94d049bb133111ebull;
}
// first `nchars` bases of a left-aligned string of WS words -> WR words (tail bits cleared)
template <int WS, int WR>
__device__ __forceinline__ void ksw_prefix(const uint32_t (&t)[WS], int nchars, uint32_t (&out)[WR]) {
  const int nb = 2 * nchars, wm = nb >> 5, rem = nb & 31;
#pragma unroll
  for (int i = 0; i < WR; ++i) {
    uint32_t v = i < WS ? t[i] : 0u;
    if (i > wm) v = 0u;
    else if (i == wm) v &= rem ? (0xffffffffu << (32 - rem)) : 0u;
    out[i] = v;
  }
}
template <int WS>
__device__ __forceinline__ void ksw_shl1(const uint32_t (&t)[WS], uint32_t (&out)[WS]) {   // drop the first base
#pragma unroll
  for (int i = 0; i < WS; ++i) out[i] = __funnelshift_l(i + 1 < WS ? t[i + 1] : 0u, t[i], 2);
}
// the four k-mers of an edge (WK words hold its k+1 bases): rec[0..1] = prefix k-mers of both strands (inserts), rec[2..3] =
// suffix k-mers (queries)
template <int WK, int WR>
__device__ __forceinline__ void ksw_edge_kmers(const uint32_t *src, int k, uint32_t (&rec)[4][WR]) {
  uint32_t fw[WK], rc[WK], sh[WK];
#pragma unroll
  for (int i = 0; i < WK; ++i) fw[i] = src[i];
  fw[WK - 1] &= 0xffffffffu << (32 * WK - 2 * (k + 1));   // drop the multiplicity if it shares the last key word
  revcomp_words<WK>(fw, k + 1, rc);
  ksw_prefix<WK, WR>(fw, k, rec[0]);
  ksw_prefix<WK, WR>(rc, k, rec[1]);
  ksw_shl1<WK>(fw, sh);
  ksw_prefix<WK, WR>(sh, k, rec[2]);
  ksw_shl1<WK>(rc, sh);
  ksw_prefix<WK, WR>(sh, k, rec[3]);
}

template <int WK, int WR>
__global__ void __launch_bounds__(kKsNT) k_ksw_hist(const uint32_t *__restrict__ edges, int64_t n_edges, int we, int k, KsGeom g,
                                                    unsigned long long *__restrict__ hist) {
  extern __shared__ uint32_t s_h[];
  const int nb = 2 * g.nslices;
  for (int i = threadIdx.x; i < nb; i += kKsNT) s_h[i] = 0;
  __syncthreads();
  for (int64_t e = (int64_t)blockIdx.x * kKsNT + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * kKsNT) {
    uint32_t rec[4][WR];
    ksw_edge_kmers<WK, WR>(edges + e * we, k, rec);
#pragma unroll
    for (int s = 0; s < 4; ++s) atomicAdd(s_h + (s >= 2 ? g.nslices : 0) + g.slice(ksw_hash<WR>(rec[s])), 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += kKsNT)
    if (s_h[i]) atomicAdd(hist + i, (unsigned long long)s_h[i]);
}

// tile = 512 edges = 2048 records of WR words, staged in bin order, copied out coalesced (word-wise)
template <int WR>
inline size_t ksw_scatter_smem_bytes(int nbins) {
  return (size_t)kKsNT * 4 * (WR * 4 + 2) + (size_t)nbins * 8 + (size_t)(nbins + 32) * 4 + 48 * 4 + 16;
}
template <int WK, int WR, int BPT>
__global__ void __launch_bounds__(kKsNT) k_ksw_scatter(const uint32_t *__restrict__ edges, int64_t n_edges, int we, int k, KsGeom g,
                                                       unsigned long long *__restrict__ cursor, uint32_t *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr int NT = kKsNT, T = NT * 4;
  const int nbins = 2 * g.nslices;
  uint32_t *stage = reinterpret_cast<uint32_t *>(smraw);                                   // [T * WR]
  long long *s_gd = reinterpret_cast<long long *>(stage + (((size_t)T * WR + 1) & ~(size_t)1));   // [nbins]
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_gd + nbins);                            // [nbins + 32]
  uint32_t *scratch = s_cnt + nbins + 32;                                                  // [48]
  uint16_t *stage_bin = reinterpret_cast<uint16_t *>(scratch + 48);                        // [T]
  const int tid = threadIdx.x;
  for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * NT + tid;
  uint32_t rec[4][WR];
  uint32_t rk[4];
  if (e < n_edges) {
    ksw_edge_kmers<WK, WR>(edges + e * we, k, rec);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint32_t b = (s >= 2 ? (uint32_t)g.nslices : 0u) + g.slice(ksw_hash<WR>(rec[s]));
      rk[s] = (b << 16) | atomicAdd(s_cnt + b, 1u);
    }
  } else {
#pragma unroll
    for (int s = 0; s < 4; ++s) rk[s] = 0xffffffffu;
  }
  __syncthreads();
  const uint32_t total = bins_scan_reserve<NT, BPT>(s_cnt, s_gd, scratch, cursor, nbins);
#pragma unroll
  for (int s = 0; s < 4; ++s)
    if (rk[s] != 0xffffffffu) {
      const uint32_t pos = s_cnt[rk[s] >> 16] + (rk[s] & 0xffffu);
#pragma unroll
      for (int c = 0; c < WR; ++c) stage[(size_t)pos * WR + c] = rec[s][c];
      stage_bin[pos] = (uint16_t)(rk[s] >> 16);
    }
  __syncthreads();
  const uint32_t total_words = total * WR;
  for (uint32_t x = tid; x < total_words; x += NT) {
    const uint32_t j = x / WR, c = x - j * WR;
    out[(s_gd[stage_bin[j]] + (long long)j) * WR + c] = stage[x];
  }
}

template <int WR>
__device__ __forceinline__ bool ksw_equal(const uint32_t *a, const uint32_t (&b)[WR]) {
  bool eq = true;
#pragma unroll
  for (int c = 0; c < WR; ++c) eq = eq && a[c] == b[c];
  return eq;
}
// inserts: tiles of 1024 records handed out in order (see k_ks_insert); slot = index of the record that claimed it
template <int WR>
__global__ void __launch_bounds__(kKsWalkNT) k_ksw_insert(const uint32_t *__restrict__ rec, int64_t n, KsGeom g, unsigned long long *table,
                                                          unsigned long long *tile_counter) {
  constexpr int NT = kKsWalkNT, R = kKsWalkR;
  __shared__ unsigned long long s_tile;
  const unsigned long long smask = (1ull << g.slice_log) - 1ull;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1ull);
    __syncthreads();
    const int64_t base = (int64_t)s_tile * (NT * R);
    __syncthreads();
    if (base >= n) return;
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
      const int64_t i = base + r * NT + threadIdx.x;
      if (i >= n) continue;
      uint32_t key[WR];
#pragma unroll
      for (int c = 0; c < WR; ++c) key[c] = rec[i * WR + c];
      unsigned long long h = g.slot(ksw_hash<WR>(key));
      const unsigned long long sbase = h & ~smask;
      for (;;) {
        const unsigned long long old = atomicCAS(table + h, kKswEmpty, (unsigned long long)i);
        if (old == kKswEmpty || ksw_equal<WR>(rec + old * WR, key)) break;
        h = sbase | ((h + 1) & smask);
      }
    }
  }
}
template <int WR>
__global__ void __launch_bounds__(kKsWalkNT) k_ksw_query(const uint32_t *__restrict__ qrec, int64_t n, const uint32_t *__restrict__ irec, KsGeom g,
                                                         const unsigned long long *__restrict__ table, uint32_t *__restrict__ miss,
                                                         unsigned long long *miss_cursor, unsigned long long *tile_counter) {
  constexpr int NT = kKsWalkNT, R = kKsWalkR;
  __shared__ unsigned long long s_tile;
  const unsigned long long smask = (1ull << g.slice_log) - 1ull;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1ull);
    __syncthreads();
    const int64_t base = (int64_t)s_tile * (NT * R);
    __syncthreads();
    if (base >= n) return;
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
      const int64_t i = base + r * NT + threadIdx.x;   // warp-uniform trip structure: every lane runs R rounds
      const bool live = i < n;
      uint32_t key[WR];
#pragma unroll
      for (int c = 0; c < WR; ++c) key[c] = live ? qrec[i * WR + c] : 0u;
      bool is_miss = false;
      if (live) {
        unsigned long long h = g.slot(ksw_hash<WR>(key));
        const unsigned long long sbase = h & ~smask;
        for (;;) {
          const unsigned long long idx = table[h];
          if (idx == kKswEmpty) { is_miss = true; break; }
          if (ksw_equal<WR>(irec + idx * WR, key)) break;
          h = sbase | ((h + 1) & smask);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, is_miss);
      if (bal) {
        unsigned long long at = 0;
        if ((threadIdx.x & 31) == 0) at = atomicAdd(miss_cursor, (unsigned long long)__popc(bal));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (is_miss) {
          uint32_t *dst = miss + (at + __popc(bal & lanemask_lt())) * WR;
#pragma unroll
          for (int c = 0; c < WR; ++c) dst[c] = key[c];
        }
      }
    }
  }
}

// items of a miss x (k bases in WR words): "$"-tail = x[1..k-1] + '$' preceded by x[0]; "$"-head = revcomp(x) preceded by '$'
template <int WR, int WI, int MODE>
__global__ void __launch_bounds__(kRangedNT) k_items_miss_w(const uint32_t *__restrict__ miss, int64_t n_miss, int k, int bin_bits, uint32_t lo,
                                                             uint32_t hi, uint32_t *__restrict__ items, unsigned long long *cursor,
                                                             unsigned long long *__restrict__ hist) {
  extern __shared__ uint32_t sh_hist[];
  const int nbins = 1 << bin_bits;
  if constexpr (MODE == 1) {
    for (int i = threadIdx.x; i < nbins; i += kRangedNT) sh_hist[i] = 0;
    __syncthreads();
  }
  for (int64_t base = (int64_t)blockIdx.x * kRangedNT; base < n_miss; base += (int64_t)gridDim.x * kRangedNT) {
    const int64_t x = base + threadIdx.x;
    uint32_t it[2][WI];
    bool in[2] = {false, false};
    if (x < n_miss) {
      uint32_t t[WR], r[WR];
#pragma unroll
      for (int c = 0; c < WR; ++c) t[c] = miss[x * WR + c];
      revcomp_words<WR>(t, k, r);
      window_item<WR, WI>(t, 1, k - 1, 0, t[0] >> 30, 0, it[0]);
      window_item<WR, WI>(r, 0, k, 1, kSentinel, 0, it[1]);
      if constexpr (MODE == 0) {
        uint32_t *dst = items + x * 2 * WI;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < WI; ++i) dst[j * WI + i] = it[j][i];
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t bin = it[j][0] >> (32 - bin_bits);
          if constexpr (MODE == 1) atomicAdd(&sh_hist[bin], 1u);
          else in[j] = bin >= lo && bin < hi;
        }
      }
    }
    if constexpr (MODE == 2) ranged_sink<WI, 2>(it, in, items, cursor);
  }
  if constexpr (MODE == 1) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += kRangedNT)
      if (sh_hist[i]) atomicAdd(hist + i, (unsigned long long)sh_hist[i]);
  }
}

}  // namespace mf
