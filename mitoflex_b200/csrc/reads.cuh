// reads.cuh -- the reads-fed partition level: canonical (k+1)-mer keys straight from 2-bit packed reads.
//
// Every thread owns 16 CONSECUTIVE base positions (one packed word), so the key of position i+1 is the key window of
// position i moved by one base: all the per-key work is funnel shifts with compile-time shift counts over W+1 words
// that are prepared once per thread (complemented words for complement(e), base-reversed and pre-aligned words for
// reverse(e)).  ~10 instructions per 64-bit key instead of ~60 for an independent extraction per position.
//
//   k_reads_hist    : digit histogram of the keys' top bits (persistent CTAs, one flush per CTA)
//   k_reads_scatter : count (shared RED) -> scan -> reserve global ranges -> keys recomputed and staged in bin order
//                     (shared atomic gives the slot) -> coalesced copy-out.  Keys are never held in registers, and
//                     unsorted keys never touch HBM.
//
// megahit semantics (KmerCounter reads the library with is_reverse=true): stored edge = reverse(e), its reverse
// complement = complement(e), key = min of the two, strand tie -> the edge itself (same bits).
#pragma once
#include "common.cuh"
#include "partition.cuh"

namespace mf {

struct ReadsSrc {
  const uint32_t *packed;   // 16 bases / word, first base in the top bits, reads back to back
  const uint32_t *sbits;    // bit g (LSB-first within word) set iff a read starts at base g
  int64_t n_bases;
  int k;
};

__device__ __forceinline__ uint32_t rev_bases(uint32_t x) {   // reverse the order of the 16 bases of a word
  x = __brev(x);
  return ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
}

template <int W>
struct KeyWindow {
  uint32_t cw[W + 1];   // complemented words of the thread's window
  uint32_t rv[W + 1];   // base-reversed window, shifted so that position i's reversed (k+1)-mer starts at base 15 - i
  uint32_t last_mask;

  __device__ __forceinline__ void init(const uint32_t *w, int K1) {
    uint32_t rr[W + 2];
#pragma unroll
    for (int j = 0; j <= W; ++j) {
      cw[j] = ~w[j];
      rr[j] = rev_bases(w[W - j]);
    }
    rr[W + 1] = 0u;
    const int q2 = 2 * (16 * W + 1 - K1);   // 2..32 bits
#pragma unroll
    for (int j = 0; j <= W; ++j) rv[j] = __funnelshift_lc(rr[j + 1], rr[j], q2);
    last_mask = 0xffffffffu << (32 * W - 2 * K1);
  }
  // first word of the canonical key of position I: min(r, c) is decided by the first differing word, so the first word
  // of the minimum is the minimum of the first words -- all a partition digit needs
  __device__ __forceinline__ uint32_t head(int I) const {
    uint32_t c0 = __funnelshift_l(cw[1], cw[0], 2 * I), r0 = __funnelshift_l(rv[1], rv[0], 2 * (15 - I));
    if constexpr (W == 1) {
      c0 &= last_mask;
      r0 &= last_mask;
    }
    return min(c0, r0);
  }
  // canonical key of the thread's position I (0..15)
  __device__ __forceinline__ void key(int I, uint32_t (&out)[W]) const {
    uint32_t c[W], r[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      c[j] = __funnelshift_l(cw[j + 1], cw[j], 2 * I);
      r[j] = __funnelshift_l(rv[j + 1], rv[j], 2 * (15 - I));
    }
    c[W - 1] &= last_mask;
    r[W - 1] &= last_mask;
    bool take_r;
    if constexpr (W == 1) {
      take_r = r[0] < c[0];
    } else if constexpr (W == 2) {
      take_r = (((unsigned long long)r[0] << 32) | r[1]) < (((unsigned long long)c[0] << 32) | c[1]);
    } else {
      take_r = false;
      bool decided = false;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        if (!decided && r[j] != c[j]) {
          take_r = r[j] < c[j];
          decided = true;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) out[j] = take_r ? r[j] : c[j];
  }
};

// bit i set iff position p0 + i (relative to the tile) starts a (k+1)-mer that lies inside one read.
// sb: the tile's start bits (bit r = relative position r); lim: positions of the tile that may start a key at all.
__device__ __forceinline__ uint32_t valid16(const uint32_t *sb, int p0, int k, int lim) {
  const int c = lim - p0;
  if (c <= 0) return 0u;
  uint32_t m = c >= 16 ? 0xffffu : ((1u << c) - 1u);
  const int first = p0 + 1, last = p0 + 15 + k;   // a read start in (j, j+k] kills position j
  for (int wi = first >> 5; wi <= (last >> 5); ++wi) {
    uint32_t x = sb[wi];
    if (wi == (first >> 5)) x &= 0xffffffffu << (first & 31);
    if (wi == (last >> 5)) x &= 0xffffffffu >> (31 - (last & 31));
    while (x) {
      const int s = wi * 32 + __ffs(x) - 1;
      x &= x - 1;
      const int lo = max(s - k, p0) - p0, hi = min(s - 1, p0 + 15) - p0;
      m &= ~(((2u << hi) - 1u) & ~((1u << lo) - 1u));
    }
  }
  return m;
}

template <int W>
struct ReadsTileCfg {
  // W = 3: 8192 keys x 12 B = 96 KB of staging, still two CTAs per SM; W = 5, 6: 4096 keys x 20 / 24 B = 80 / 96 KB, likewise
  static constexpr int NT = W <= 3 ? 512 : (W <= 6 ? 256 : 128);
  static constexpr int T = NT * 16;   // base positions per tile
};
__host__ __device__ inline int reads_seq_words(int NT, int W) { return NT + W + 1; }
__host__ __device__ inline int reads_bit_words(int NT, int k) { return NT / 2 + (k + 15 + 31) / 32 + 2; }

// loads the tile's packed words and start bits; returns the number of positions that may start a key
template <int W, int NT>
__device__ __forceinline__ int reads_load_tile(const ReadsSrc &src, int64_t tile, uint32_t *seq, uint32_t *sb) {
  constexpr int T = NT * 16;
  const int64_t base = tile * (int64_t)T;
  const int64_t total_words = (src.n_bases + 15) >> 4, w0 = base >> 4;
  const int nsw = reads_seq_words(NT, W);
  for (int i = threadIdx.x; i < nsw; i += NT) seq[i] = (w0 + i < total_words) ? src.packed[w0 + i] : 0u;
  const int64_t total_bw = (src.n_bases + 31) >> 5, b0 = base >> 5;
  const int nbw = reads_bit_words(NT, src.k);
  for (int i = threadIdx.x; i < nbw; i += NT) sb[i] = (b0 + i < total_bw) ? src.sbits[b0 + i] : 0u;
  const int64_t lim = src.n_bases - (src.k + 1) + 1 - base;   // positions with a whole (k+1)-mer before the end
  return (int)(lim < 0 ? 0 : (lim > T ? T : lim));
}

// ============================================================ histogram
// invalid positions add to one of 32 per-lane dummy counters behind the bins, so the loop has no branch
template <int W, int NT, bool RANGED>
__global__ void __launch_bounds__(NT) k_reads_hist(ReadsSrc src, LevelArgs a, unsigned long long *__restrict__ hist,
                                                   int64_t tile0, int64_t ntiles, int64_t tstride) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int nbins = 1 << a.nbits;
  uint32_t *s_hist = smem;                   // [nbins + 32]
  uint32_t *seq = s_hist + nbins + 32;
  uint32_t *sb = seq + reads_seq_words(NT, W);
  const int tid = threadIdx.x;
  const int dsh = 32 - a.nbits;
  const uint32_t dlo = a.dlo, dspan = a.dhi - a.dlo;
  const uint32_t dummy = (uint32_t)nbins + (tid & 31);
  for (int i = tid; i < nbins + 32; i += NT) s_hist[i] = 0;
  for (int64_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {   // tiles tile0, tile0 + tstride, ... (a strided sample when tstride > 1)
    const int64_t tile = tile0 + ti * tstride;
    __syncthreads();
    const int lim = reads_load_tile<W, NT>(src, tile, seq, sb);
    __syncthreads();
    const uint32_t vm = valid16(sb, tid * 16, src.k, lim);
    if (vm == 0u) continue;
    KeyWindow<W> kw;
    kw.init(seq + tid, src.k + 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t d = kw.head(i) >> dsh;
      bool ok = (vm >> i) & 1u;
      if constexpr (RANGED) ok = ok && (d - dlo) < dspan;
      atomicAdd(s_hist + (ok ? d : dummy), 1u);
    }
  }
  __syncthreads();
  for (int b = tid; b < nbins; b += NT) {
    const uint32_t c = s_hist[b];
    if (c) atomicAdd(hist + b, (unsigned long long)c);
  }
}

// ============================================================ scatter
// dynamic smem (uint32 units): s_cnt[nbins+32] | pad | s_gd i64[nbins] | scratch[36] | stage[T*W] | seq | sb
template <int W>
__host__ __device__ inline size_t reads_scatter_smem_bytes(int NT, int nbits, int k) {
  const size_t nb = (size_t)1 << nbits;
  return (nb + 32 + 2 + 2 * nb + 36 + 2 + (size_t)NT * 16 * W + reads_seq_words(NT, W) + reads_bit_words(NT, k)) * 4;
}

template <int W, int NT, int BPT, bool RANGED>
__global__ void __launch_bounds__(NT) k_reads_scatter(ReadsSrc src, LevelArgs a, unsigned long long *__restrict__ cursor,
                                                      uint32_t *__restrict__ out, int64_t tile0) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int T = NT * 16;
  const int nbins = 1 << a.nbits;
  const int tid = threadIdx.x;
  uint32_t *s_cnt = smem;                      // [nbins + 32]: bins, then per-lane dummies for dropped positions
  uint32_t *s_gd32 = s_cnt + nbins + 32;
  if ((reinterpret_cast<uintptr_t>(s_gd32) & 7) != 0) s_gd32 += 1;
  long long *s_gd = reinterpret_cast<long long *>(s_gd32);
  uint32_t *scratch = s_gd32 + 2 * nbins;
  uint32_t *stage = scratch + 36;
  if ((reinterpret_cast<uintptr_t>(stage) & 7) != 0) stage += 1;
  uint32_t *seq = stage + (size_t)T * W;
  uint32_t *sb = seq + reads_seq_words(NT, W);
  const int dsh = 32 - a.nbits;
  const uint32_t dlo = a.dlo, dspan = a.dhi - a.dlo;
  const uint32_t dummy = (uint32_t)nbins + (tid & 31);

  for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
  const int lim = reads_load_tile<W, NT>(src, tile0 + blockIdx.x, seq, sb);
  __syncthreads();
  const uint32_t vm = valid16(sb, tid * 16, src.k, lim);
  KeyWindow<W> kw;
  kw.init(seq + tid, src.k + 1);
  // phase A: count (only the first key word matters for the digit)
  if (vm) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t d = kw.head(i) >> dsh;
      bool ok = (vm >> i) & 1u;
      if constexpr (RANGED) ok = ok && (d - dlo) < dspan;
      atomicAdd(s_cnt + (ok ? d : dummy), 1u);
    }
  }
  __syncthreads();
  const uint32_t total = bins_scan_reserve<NT, BPT>(s_cnt, s_gd, scratch, cursor, nbins, a.limit);
  // phase B: keys again, each takes the next free slot of its bin (the partition need not be stable)
  if (vm) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t key[W];
      kw.key(i, key);
      const uint32_t d = key[0] >> dsh;
      bool ok = (vm >> i) & 1u;
      if constexpr (RANGED) ok = ok && (d - dlo) < dspan;
      const uint32_t pos = atomicAdd(s_cnt + (ok ? d : dummy), 1u);
      if (ok) {
        if constexpr (W == 2) {
          *reinterpret_cast<uint2 *>(stage + (size_t)pos * 2) = make_uint2(key[0], key[1]);
        } else {
#pragma unroll
          for (int c = 0; c < W; ++c) stage[(size_t)pos * W + c] = key[c];
        }
      }
    }
  }
  __syncthreads();
  // coalesced copy-out: consecutive staged records of a bin go to consecutive global records
  if constexpr (W == 2) {
    const uint2 *st2 = reinterpret_cast<const uint2 *>(stage);
    uint2 *out2 = reinterpret_cast<uint2 *>(out);
    if (a.bin_base) {   // fused exchange: bin b lives at its own (possibly peer-GPU) address
      for (uint32_t j = tid; j < total; j += NT) {
        const uint2 v = st2[j];
        const uint32_t d = v.x >> dsh;
        reinterpret_cast<uint2 *>(a.bin_base[d])[s_gd[d] + (long long)j] = v;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) {   // a tile stages at most 16 keys per thread: unrolled, all shared loads in flight
        const uint32_t j = (uint32_t)(q * NT + tid);
        if (j < total) {
          const uint2 v = st2[j];
          const long long gd = s_gd[v.x >> dsh];
          if (gd != kDropRun) out2[gd + (long long)j] = v;
        }
      }
    }
  } else {
    const uint32_t total_words = total * W;
    for (uint32_t x = tid; x < total_words; x += NT) {
      const uint32_t j = x / W, c = x - j * W;
      const uint32_t d = stage[(size_t)j * W] >> dsh;
      uint32_t *dst = a.bin_base ? reinterpret_cast<uint32_t *>(a.bin_base[d]) : out;
      const long long gd = s_gd[d];
      if (gd != kDropRun) dst[(gd + (long long)j) * W + c] = stage[x];
    }
  }
}

}  // namespace mf
