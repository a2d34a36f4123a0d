// kmerset.cuh -- which "$" dummies of an edge set reach the graph (the item filter of seq2sdbg / read2sdbg), second generation.
//
// Of the 6 items SeqToSdbg generates per edge (both strands: "$"-head, real item, "$"-tail) only the 2 real ones always
// survive Lv2Postprocess; a "$"-tail item of a k-mer x survives only if x has no OUTGOING solid edge, a "$"-head item of
// revcomp(x) only under the same condition (its k-mer has no incoming edge).  Both are membership tests in
//     PS = { first k bases of t : t in (edges U revcomp(edges)) },
// asked for every suffix k-mer of the same strings.  A query needs no payload: both dummies follow from x alone, so
//     misses = { x in suffix k-mers : x not in PS }          (dead ends of the graph; a handful per thousand edges)
// is all the generator needs beside the edges.  Dropping the other dummies changes no output (argument in items.cuh).
//
// Round 1 probed a 4.3 GB open-addressing table in HBM straight from the (sorted) edges: half of the 4 accesses per edge went
// to random DRAM rows, ~130 bytes fetched per 8-byte probe, 31 GB of traffic for 5 GB of algorithmic bytes (ncu r1g/r1i).  Here
// the table is cut into SLICES of 16 MB by the top bits of a hash, all inserts and all queries are first scattered by slice
// (one histogram pass + one staged, coalesced scatter pass over the edges), and the insert / query kernels walk the records
// in slice order: the ~300 CTAs in flight work inside one or two slices at a time, so a table line is fetched from DRAM once
// and then served by L2 for all of its ~10 accesses.  Hash slices are uniform whatever the genome's composition; slots and
// indices are 64-bit, so the edge count is bounded by memory only; k <= 31 uses 8-byte keys, k <= 63 16-byte keys (128-bit CAS).
#pragma once
#include "common.cuh"
#include "items.cuh"
#include "partition.cuh"

namespace mf {

constexpr int kKsSliceLog = 21;     // slots per slice: 2^21 x 8 B = 16 MB (x 16 B = 32 MB for k >= 32)
constexpr int kKsMaxSlices = 1024;  // beyond that the slices grow instead (scatter bins = 2 x slices)
constexpr int kKsNT = 512;

template <int KW>
struct KsKey;
template <>
struct KsKey<1> {   // k <= 31: up to 64 bits, left aligned
  unsigned long long v;
  using Slot = unsigned long long;
  __device__ __forceinline__ static KsKey load_edge(const uint32_t *src, int wk, int k) {
    unsigned long long t = (unsigned long long)src[0] << 32;
    if (wk == 2) t |= src[1];
    return KsKey{t & (~0ull << (64 - 2 * (k + 1)))};   // drop the multiplicity if it shares the last key word
  }
  __device__ __forceinline__ KsKey revcomp(int nchars) const { return KsKey{revcomp64(v, nchars)}; }
  __device__ __forceinline__ KsKey prefix(int nchars) const { return KsKey{v & (~0ull << (64 - 2 * nchars))}; }
  __device__ __forceinline__ KsKey shl_chars(int c) const { return KsKey{v << (2 * c)}; }
  __device__ __forceinline__ unsigned long long hash() const {
    unsigned long long x = v * 0x9e3779b97f4a7c15ull;
    x ^= x >> 32;
    return x * 0xbf58476d1ce4e5b9ull;
  }
  __device__ __forceinline__ Slot pack() const { return v; }
  __device__ __forceinline__ static Slot empty() { return ~0ull; }   // 2k <= 62 bits used: never a k-mer
  __device__ __forceinline__ void words(uint32_t (&t)[2]) const { t[0] = (uint32_t)(v >> 32); t[1] = (uint32_t)v; }
  __device__ __forceinline__ uint32_t first_char() const { return (uint32_t)(v >> 62); }
};
template <>
struct KsKey<2> {   // 32 <= k <= 63: up to 128 bits
  K128 v;
  using Slot = unsigned __int128;
  __device__ __forceinline__ static KsKey load_edge(const uint32_t *src, int wk, int k) {
    K128 t;
    t.hi = ((unsigned long long)src[0] << 32) | src[1];
    t.lo = (unsigned long long)src[2] << 32;
    if (wk >= 4) t.lo |= src[3];
    return KsKey{mask_top128(t, 2 * (k + 1))};
  }
  __device__ __forceinline__ KsKey revcomp(int nchars) const { return KsKey{revcomp128(v, nchars)}; }
  __device__ __forceinline__ KsKey prefix(int nchars) const { return KsKey{mask_top128(v, 2 * nchars)}; }
  __device__ __forceinline__ KsKey shl_chars(int c) const { return KsKey{shl128(v, 2 * c)}; }
  __device__ __forceinline__ unsigned long long hash() const {
    unsigned long long x = (v.hi ^ (v.lo * 0x94d049bb133111ebull)) * 0x9e3779b97f4a7c15ull;
    x ^= x >> 32;
    return x * 0xbf58476d1ce4e5b9ull;
  }
  __device__ __forceinline__ Slot pack() const { return pack128(v); }
  __device__ __forceinline__ static Slot empty() { return ~(unsigned __int128)0; }   // 2k <= 126 bits used
  __device__ __forceinline__ void words(uint32_t (&t)[4]) const {
    t[0] = (uint32_t)(v.hi >> 32); t[1] = (uint32_t)v.hi; t[2] = (uint32_t)(v.lo >> 32); t[3] = (uint32_t)v.lo;
  }
  __device__ __forceinline__ uint32_t first_char() const { return (uint32_t)(v.hi >> 62); }
};

// the four k-mers of an edge: inserts = prefix k-mers of both strands, queries = suffix k-mers of both strands
template <int KW>
__device__ __forceinline__ void ks_edge_kmers(const uint32_t *src, int wk, int k, KsKey<KW> (&ins)[2], KsKey<KW> (&qry)[2]) {
  const KsKey<KW> fw = KsKey<KW>::load_edge(src, wk, k);
  const KsKey<KW> rc = fw.revcomp(k + 1);
  ins[0] = fw.prefix(k);
  ins[1] = rc.prefix(k);
  qry[0] = fw.shl_chars(1).prefix(k);
  qry[1] = rc.shl_chars(1).prefix(k);
}
struct KsGeom {
  int log_slots;    // table slots = 2^log_slots
  int slice_log;    // slots per slice
  int nslices;      // 2^(log_slots - slice_log), at least 1
  __device__ __forceinline__ unsigned long long slot(unsigned long long h) const { return h >> (64 - log_slots); }
  __device__ __forceinline__ uint32_t slice(unsigned long long h) const { return (uint32_t)(h >> (64 - log_slots) >> slice_log); }
};

// ---- pass 1: records per (kind, slice): hist[kind * nslices + slice]
template <int KW>
__global__ void __launch_bounds__(kKsNT) k_ks_hist(const uint32_t *__restrict__ edges, int64_t n_edges, int wk, int we, int k, KsGeom g,
                                                   unsigned long long *__restrict__ hist) {
  extern __shared__ uint32_t s_h[];
  const int nb = 2 * g.nslices;
  for (int i = threadIdx.x; i < nb; i += kKsNT) s_h[i] = 0;
  __syncthreads();
  for (int64_t e = (int64_t)blockIdx.x * kKsNT + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * kKsNT) {
    KsKey<KW> ins[2], qry[2];
    ks_edge_kmers<KW>(edges + e * we, wk, k, ins, qry);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      atomicAdd(s_h + g.slice(ins[s].hash()), 1u);
      atomicAdd(s_h + g.nslices + g.slice(qry[s].hash()), 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += kKsNT)
    if (s_h[i]) atomicAdd(hist + i, (unsigned long long)s_h[i]);
}

// ---- pass 2: scatter the 4 k-mers of every edge by (kind, slice); tile = NT * EPT edges, staged, copied out coalesced
template <int KW>
struct KsScatterCfg {
  static constexpr int EPT = KW == 1 ? 4 : 2;          // edges per thread
  static constexpr int T = kKsNT * EPT * 4;            // records per tile
};
template <int KW>
inline size_t ks_scatter_smem_bytes(int nbins) {
  return (size_t)KsScatterCfg<KW>::T * (8 * KW + 2) + (size_t)nbins * 8 + (size_t)(nbins + 32) * 4 + 48 * 4 + 16;
}
template <int KW, int BPT>
__global__ void __launch_bounds__(kKsNT) k_ks_scatter(const uint32_t *__restrict__ edges, int64_t n_edges, int wk, int we, int k, KsGeom g,
                                                      unsigned long long *__restrict__ cursor, typename KsKey<KW>::Slot *__restrict__ out,
                                                      const unsigned long long *__restrict__ bin_base = nullptr) {
  extern __shared__ __align__(128) unsigned char smraw[];
  using C = KsScatterCfg<KW>;
  using Slot = typename KsKey<KW>::Slot;
  constexpr int NT = kKsNT, EPT = C::EPT, T = C::T;
  const int nbins = 2 * g.nslices;
  Slot *stage = reinterpret_cast<Slot *>(smraw);                         // [T]
  long long *s_gd = reinterpret_cast<long long *>(stage + T);            // [nbins]
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_gd + nbins);          // [nbins + 32]
  uint32_t *scratch = s_cnt + nbins + 32;                                // [48]
  uint16_t *stage_bin = reinterpret_cast<uint16_t *>(scratch + 48);      // [T]
  const int tid = threadIdx.x;
  for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
  __syncthreads();
  const int64_t e0 = (int64_t)blockIdx.x * (NT * EPT);
  Slot rec[EPT][4];
  uint32_t rk[EPT][4];   // bin << 16 | rank within the bin
#pragma unroll
  for (int q = 0; q < EPT; ++q) {
    const int64_t e = e0 + q * NT + tid;
    if (e < n_edges) {
      KsKey<KW> ins[2], qry[2];
      ks_edge_kmers<KW>(edges + e * we, wk, k, ins, qry);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const uint32_t bi = g.slice(ins[s].hash()), bq = (uint32_t)g.nslices + g.slice(qry[s].hash());
        rec[q][s] = ins[s].pack();
        rec[q][2 + s] = qry[s].pack();
        rk[q][s] = (bi << 16) | atomicAdd(s_cnt + bi, 1u);
        rk[q][2 + s] = (bq << 16) | atomicAdd(s_cnt + bq, 1u);
      }
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s) rk[q][s] = 0xffffffffu;
    }
  }
  __syncthreads();
  const uint32_t total = bins_scan_reserve<NT, BPT>(s_cnt, s_gd, scratch, cursor, nbins);
#pragma unroll
  for (int q = 0; q < EPT; ++q)
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if (rk[q][s] != 0xffffffffu) {
        const uint32_t pos = s_cnt[rk[q][s] >> 16] + (rk[q][s] & 0xffffu);
        stage[pos] = rec[q][s];
        stage_bin[pos] = (uint16_t)(rk[q][s] >> 16);
      }
  __syncthreads();
  // copy-out: consecutive staged records of a bin go to consecutive global records
  // (bin_base: the multi-GPU filter of ksdist.cu -- bin b lives at its own, possibly peer-GPU, address and `cursor` counts from 0)
  for (uint32_t j = tid; j < total; j += NT) {
    const uint32_t b = stage_bin[j];
    Slot *dst = bin_base ? reinterpret_cast<Slot *>(bin_base[b]) : out;
    dst[s_gd[b] + (long long)j] = stage[j];
  }
}

// ---- pass 3: inserts, in slice order.  The records of a slice are contiguous, and tiles of 1024 records are handed out IN ORDER by
// an atomic counter, so the CTAs in flight (8 per SM) always work on the ~1.2 M most recent records = one or two slices, whose
// table lines stay in L2 (tiles per WARP and 8 records per thread were both slower: 3.3 / 4.8 ms against 2.06 ms on the 1.8 Gbp
// sample, ncu r2h/r2i).  (A plain grid-stride loop does not keep that window: after a few hundred dependent DRAM round trips
// per thread the fast warps are dozens of slices ahead of the slow ones -- ncu r2d: 131 bytes of DRAM read per insert.)
constexpr int kKsWalkNT = 256, kKsWalkR = 4;
template <int KW>
__device__ __forceinline__ KsKey<KW> ks_unpack(typename KsKey<KW>::Slot x) {
  KsKey<KW> kx;
  if constexpr (KW == 1) kx.v = x; else { kx.v.hi = (unsigned long long)(x >> 64); kx.v.lo = (unsigned long long)x; }
  return kx;
}
template <int KW>
__global__ void __launch_bounds__(kKsWalkNT) k_ks_insert(const typename KsKey<KW>::Slot *__restrict__ rec, int64_t n, KsGeom g,
                                                         typename KsKey<KW>::Slot *table, unsigned long long *tile_counter) {
  using Slot = typename KsKey<KW>::Slot;
  constexpr int NT = kKsWalkNT, R = kKsWalkR;
  __shared__ unsigned long long s_tile;
  const Slot empty = KsKey<KW>::empty();
  const unsigned long long smask = (1ull << g.slice_log) - 1ull;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1ull);
    __syncthreads();
    const int64_t base = (int64_t)s_tile * (NT * R);
    __syncthreads();
    if (base >= n) return;
    Slot x[R], c[R];
    unsigned long long h[R];
    bool live[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t i = base + r * NT + threadIdx.x;
      live[r] = i < n;
      x[r] = rec[live[r] ? i : n - 1];
      h[r] = g.slot(ks_unpack<KW>(x[r]).hash());
    }
#pragma unroll
    for (int r = 0; r < R; ++r) c[r] = live[r] ? atomicCAS(table + h[r], empty, x[r]) : empty;   // all first probes in flight together
#pragma unroll
    for (int r = 0; r < R; ++r) {
      // linear probing inside the slice (wrapping at its end): a probe sequence never leaves the slice's L2-resident lines
      const unsigned long long sbase = h[r] & ~smask;
      while (c[r] != empty && c[r] != x[r]) {
        h[r] = sbase | ((h[r] + 1) & smask);
        c[r] = atomicCAS(table + h[r], empty, x[r]);
      }
    }
  }
}
// ---- pass 4: queries, in slice order (same walk); a k-mer that is not in the set goes to the miss list (warp-aggregated append)
template <int KW>
__global__ void __launch_bounds__(kKsWalkNT) k_ks_query(const typename KsKey<KW>::Slot *__restrict__ rec, int64_t n, KsGeom g,
                                                        const typename KsKey<KW>::Slot *__restrict__ table,
                                                        typename KsKey<KW>::Slot *__restrict__ miss, unsigned long long *miss_cursor,
                                                        unsigned long long *tile_counter) {
  using Slot = typename KsKey<KW>::Slot;
  constexpr int NT = kKsWalkNT, R = kKsWalkR;
  __shared__ unsigned long long s_tile;
  const Slot empty = KsKey<KW>::empty();
  const unsigned long long smask = (1ull << g.slice_log) - 1ull;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1ull);
    __syncthreads();
    const int64_t base = (int64_t)s_tile * (NT * R);
    __syncthreads();
    if (base >= n) return;
    Slot x[R], c[R];
    unsigned long long h[R];
    bool live[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t i = base + r * NT + threadIdx.x;
      live[r] = i < n;
      x[r] = rec[live[r] ? i : n - 1];
      h[r] = g.slot(ks_unpack<KW>(x[r]).hash());
    }
#pragma unroll
    for (int r = 0; r < R; ++r) c[r] = table[h[r]];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const unsigned long long sbase = h[r] & ~smask;
      while (c[r] != empty && c[r] != x[r]) {
        h[r] = sbase | ((h[r] + 1) & smask);
        c[r] = table[h[r]];
      }
      const bool is_miss = live[r] && c[r] == empty;
      const unsigned bal = __ballot_sync(0xffffffffu, is_miss);
      if (bal) {
        unsigned long long at = 0;
        if ((threadIdx.x & 31) == 0) at = atomicAdd(miss_cursor, (unsigned long long)__popc(bal));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (is_miss) miss[at + __popc(bal & lanemask_lt())] = x[r];
      }
    }
  }
}

// ---- item generation from the edges (2 real items each) and from the miss list (2 dummies each).
// MODE 0: real item j of edge e at items[2e + j]; dummies of miss i at items[dummy_base + 2i + j].
// MODE 1: nothing written, per-bin counts of the items' top bin_bits bits -> hist (sdbg rounds).
// MODE 2: only the items whose bin lies in [lo, hi), appended through a warp-aggregated cursor.
template <int WK, int WE, int WI, int MODE>
__global__ void __launch_bounds__(kRangedNT) k_items_real(const uint32_t *__restrict__ edges, int64_t n_edges, int k, int bin_bits, uint32_t lo,
                                                           uint32_t hi, uint32_t *__restrict__ items, unsigned long long *cursor,
                                                           unsigned long long *__restrict__ hist) {
  extern __shared__ uint32_t sh_hist[];
  const int nbins = 1 << bin_bits;
  if constexpr (MODE == 1) {
    for (int i = threadIdx.x; i < nbins; i += kRangedNT) sh_hist[i] = 0;
    __syncthreads();
  }
  for (int64_t base = (int64_t)blockIdx.x * kRangedNT; base < n_edges; base += (int64_t)gridDim.x * kRangedNT) {
    const int64_t e = base + threadIdx.x;
    const bool live = e < n_edges;
    uint32_t it[2][WI];
    bool in[2] = {false, false};
    if (live) {
      uint32_t fw[WK], rc[WK];
      const uint32_t *src = edges + e * WE;
#pragma unroll
      for (int i = 0; i < WK; ++i) fw[i] = src[i];
      const uint32_t mult = src[WE - 1] & 0xffffu;
      fw[WK - 1] &= 0xffffffffu << (32 * WK - 2 * (k + 1));   // drop the multiplicity if it shares the last key word
      revcomp_words<WK>(fw, k + 1, rc);
      window_item<WK, WI>(fw, 1, k, 1, fw[0] >> 30, mult, it[0]);
      window_item<WK, WI>(rc, 1, k, 1, rc[0] >> 30, mult, it[1]);
      if constexpr (MODE == 0) {
        uint32_t *dst = items + e * 2 * WI;
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
// a miss x (k bases): "$"-tail item = x[1..k-1] + '$' with preceding base x[0]; "$"-head item = revcomp(x) preceded by '$'
template <int KW, int WI, int MODE>
__global__ void __launch_bounds__(kRangedNT) k_items_miss(const typename KsKey<KW>::Slot *__restrict__ miss, int64_t n_miss, int k, int bin_bits,
                                                           uint32_t lo, uint32_t hi, uint32_t *__restrict__ items, unsigned long long *cursor,
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
      const typename KsKey<KW>::Slot raw = miss[x];
      KsKey<KW> kx;
      if constexpr (KW == 1) kx.v = raw; else { kx.v.hi = (unsigned long long)(raw >> 64); kx.v.lo = (unsigned long long)raw; }
      uint32_t t[2 * KW], r[2 * KW];
      kx.words(t);
      kx.revcomp(k).words(r);
      window_item<2 * KW, WI>(t, 1, k - 1, 0, kx.first_char(), 0, it[0]);
      window_item<2 * KW, WI>(r, 0, k, 1, kSentinel, 0, it[1]);
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
