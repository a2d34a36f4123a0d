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
  stage += ((16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15) >> 2;   // 16-byte aligned (staged_copy_out)
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
          stage_store<W>(stage + (size_t)pos * W, key);
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
    staged_copy_out<W, NT>(stage, s_gd, total, a, out, [dsh](const uint32_t *r) { return r[0] >> dsh; });
  }
}

// ============================================================ scatter, compacted: wide keys on short reads
// k_reads_scatter spends its instructions per POSITION (16 per thread, predicated); with k = 141 on 150-base reads 6 % of the
// positions start a key, at k = 119 20 %.  Here the valid positions of a tile are listed first (every thread appends the set
// bits of its valid16 mask behind a block scan) and the threads then walk that LIST: each key is extracted on its own from
// the tile's words in shared memory -- no rolling window, ~4W + 25 instructions per key instead of ~2W + 8 per position --
// counted, staged in bin order and copied out like above, in batches of KCAP keys so that the staging area does not have to
// hold a key per position.  Tiles are 8192 positions whatever the width, so the few keys of a tile still make runs.
template <int W>
struct ReadsCompactCfg {
  static constexpr int NT = 512;
  static constexpr int T = NT * 16;
  static constexpr int KCAP = W <= 4 ? 4096 : (W <= 6 ? 2560 : (W <= 8 ? 2048 : 1536));   // 51-65 KB of staging
};
// dynamic smem (uint32 units): s_cnt[nbins+32] | pad | s_gd i64[nbins] | scratch[36] | s_pc[NT] | s_pos u16[T] | pad | stage[KCAP*W] | seq | sb
template <int W>
__host__ __device__ inline size_t reads_compact_smem_bytes(int nbits, int k) {
  using C = ReadsCompactCfg<W>;
  const size_t nb = (size_t)1 << nbits;
  return (nb + 32 + 2 + 2 * nb + 36 + C::NT + C::T / 2 + 2 + (size_t)C::KCAP * W + reads_seq_words(C::NT, W) + reads_bit_words(C::NT, k)) * 4;
}
// first word of the canonical key that starts at tile position p (K1 = k + 1 > 16)
__device__ __forceinline__ uint32_t reads_head_at(const uint32_t *seq, int p, int K1) {
  const int pw = p >> 4, q = p + K1 - 16, qw = q >> 4;
  const uint32_t c0 = ~__funnelshift_l(seq[pw + 1], seq[pw], 2 * (p & 15));
  const uint32_t r0 = rev_bases(__funnelshift_l(seq[qw + 1], seq[qw], 2 * (q & 15)));
  return min(c0, r0);
}
// the canonical key that starts at tile position p: min(complement(e), reverse(e)), W words, big-endian (= KeyWindow::key).
// Both strands come from the W forward words f of the window: reversing the bases of the whole 16 W-base string puts the
// 16 W - K1 pad bases (zeros) in front, so reverse(e) = (rev_bases(f[W-1]), ..., rev_bases(f[0])) shifted left by the pad.
template <int W>
__device__ __forceinline__ void reads_key_at(const uint32_t *seq, int p, int K1, uint32_t (&out)[W]) {
  const int pw = p >> 4, ps = 2 * (p & 15);
  const int pad2 = 2 * (16 * W - K1);   // 0..30 bits
  uint32_t wnd[W + 1], f[W], c[W], rr[W + 1], r[W];
#pragma unroll
  for (int j = 0; j <= W; ++j) wnd[j] = seq[pw + j];
#pragma unroll
  for (int j = 0; j < W; ++j) f[j] = __funnelshift_l(wnd[j + 1], wnd[j], ps);
  const uint32_t last_mask = 0xffffffffu << pad2;
  f[W - 1] &= last_mask;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    c[j] = ~f[j];
    rr[j] = rev_bases(f[W - 1 - j]);
  }
  c[W - 1] &= last_mask;
  rr[W] = 0u;
#pragma unroll
  for (int j = 0; j < W; ++j) r[j] = __funnelshift_l(rr[j + 1], rr[j], pad2);
  bool take_r = false, decided = false;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (!decided && r[j] != c[j]) {
      take_r = r[j] < c[j];
      decided = true;
    }
  }
#pragma unroll
  for (int j = 0; j < W; ++j) out[j] = take_r ? r[j] : c[j];
}

// positions [pos0, pos1) of the reads (pos0 a multiple of 16); one tile of 8192 positions per CTA
template <int W>
__global__ void __launch_bounds__(ReadsCompactCfg<W>::NT, 2) k_reads_scatter_compact(ReadsSrc src, LevelArgs a, unsigned long long *__restrict__ cursor,
                                                                                   uint32_t *__restrict__ out, int64_t pos0, int64_t pos1) {
  extern __shared__ __align__(16) uint32_t smem[];
  using C = ReadsCompactCfg<W>;
  constexpr int NT = C::NT, T = C::T, KCAP = C::KCAP;
  const int nbins = 1 << a.nbits, tid = threadIdx.x, K1 = src.k + 1;
  uint32_t *s_cnt = smem;
  uint32_t *s_gd32 = s_cnt + nbins + 32;
  if ((reinterpret_cast<uintptr_t>(s_gd32) & 7) != 0) s_gd32 += 1;
  long long *s_gd = reinterpret_cast<long long *>(s_gd32);
  uint32_t *scratch = s_gd32 + 2 * nbins;
  uint32_t *s_pc = scratch + 36;
  uint16_t *s_pos = reinterpret_cast<uint16_t *>(s_pc + NT);
  uint32_t *stage = s_pc + NT + T / 2;
  stage += ((16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15) >> 2;   // 16-byte aligned (staged_copy_out)
  uint32_t *seq = stage + (size_t)KCAP * W;
  uint32_t *sb = seq + reads_seq_words(NT, W);
  const int dsh = 32 - a.nbits;
  const uint32_t dlo = a.dlo, dspan = a.dhi - a.dlo;
  const bool ranged = dlo != 0u || a.dhi != (1u << a.nbits);
  const uint32_t dummy = (uint32_t)nbins + (tid & 31);

  // the tile's words and start bits (positions relative to `base`); keys may start below `lim` only
  const int64_t base = pos0 + (int64_t)blockIdx.x * T;
  {
    const int64_t total_words = (src.n_bases + 15) >> 4, w0 = base >> 4;
    const int nsw = reads_seq_words(NT, W);
    for (int i = tid; i < nsw; i += NT) seq[i] = (w0 + i < total_words) ? src.packed[w0 + i] : 0u;
    const int64_t total_bw = (src.n_bases + 31) >> 5, b0 = base >> 5;
    const int nbw = reads_bit_words(NT, src.k);
    for (int i = tid; i < nbw; i += NT) sb[i] = (b0 + i < total_bw) ? src.sbits[b0 + i] : 0u;
  }
  int64_t lim64 = src.n_bases - K1 + 1 - base;
  if (pos1 - base < lim64) lim64 = pos1 - base;
  const int lim = (int)(lim64 < 0 ? 0 : (lim64 > T ? T : lim64));
  __syncthreads();
  const uint32_t vm = valid16(sb, tid * 16, src.k, lim);
  s_pc[tid] = __popc(vm);
  __syncthreads();
  const uint32_t V = block_excl_scan<NT>(s_pc, NT, scratch);
  {
    uint32_t at = s_pc[tid];
    for (uint32_t m = vm; m; m &= m - 1) s_pos[at++] = (uint16_t)(tid * 16 + __ffs(m) - 1);
  }
  __syncthreads();
  for (uint32_t j0 = 0; j0 < V; j0 += KCAP) {
    const uint32_t nb = min((uint32_t)KCAP, V - j0);
    for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
    __syncthreads();
    for (uint32_t j = tid; j < nb; j += NT) {
      const uint32_t d = reads_head_at(seq, s_pos[j0 + j], K1) >> dsh;
      const bool ok = !ranged || (d - dlo) < dspan;
      atomicAdd(s_cnt + (ok ? d : dummy), 1u);
    }
    __syncthreads();
    const uint32_t total = bins_scan_reserve<NT, 2>(s_cnt, s_gd, scratch, cursor, nbins, a.limit);   // up to 2 NT = 1024 bins
    for (uint32_t j = tid; j < nb; j += NT) {
      uint32_t key[W];
      reads_key_at<W>(seq, s_pos[j0 + j], K1, key);
      const uint32_t d = key[0] >> dsh;
      const bool ok = !ranged || (d - dlo) < dspan;
      const uint32_t pos = atomicAdd(s_cnt + (ok ? d : dummy), 1u);
      if (ok) {
        stage_store<W>(stage + (size_t)pos * W, key);
      }
    }
    __syncthreads();
    staged_copy_out<W, NT>(stage, s_gd, total, a, out, [dsh](const uint32_t *r) { return r[0] >> dsh; });
    __syncthreads();
  }
}

// ---- the same, with the batch cut loose from the listing pass: a CTA owns 32768 positions (one load of the words and start
// bits), lists the valid ones 8192 positions at a time and runs a batch whenever KCAP listed positions have come together -- at
// k = 141 a listing pass finds ~500 keys, and a batch per pass means the block-wide scan, the global reservation and their
// barriers are paid for 500 keys instead of 1536.
template <int W>
struct ReadsCompact2Cfg {
  static constexpr int NT = 512;
  static constexpr int SUB = NT * 16;     // positions per listing pass
  static constexpr int NSUB = 4;
  static constexpr int T = SUB * NSUB;    // positions per CTA (< 65536: listed as uint16)
  static constexpr int KCAP = W <= 4 ? 3584 : ReadsCompactCfg<W>::KCAP;
  static constexpr int PCAP = KCAP + SUB;   // listed positions held at once: what a batch left over + one pass
};
// dynamic smem (uint32 units): s_cnt[nbins+32] | pad | s_gd i64[nbins] | scratch[36] | s_pc[NT] | s_pos u16[PCAP] | pad | stage[KCAP*W] | seq | sb
template <int W>
__host__ __device__ inline int reads_compact2_seq_words() { return ReadsCompact2Cfg<W>::T / 16 + W + 1; }
__host__ __device__ inline int reads_compact2_bit_words(int T, int k) { return T / 32 + (k + 15 + 31) / 32 + 2; }
template <int W>
__host__ __device__ inline size_t reads_compact2_smem_bytes(int nbits, int k) {
  using C = ReadsCompact2Cfg<W>;
  const size_t nb = (size_t)1 << nbits;
  return (nb + 32 + 2 + 2 * nb + 36 + C::NT + C::PCAP / 2 + 2 + (size_t)C::KCAP * W + reads_compact2_seq_words<W>() + reads_compact2_bit_words(C::T, k)) * 4;
}

template <int W>
__global__ void __launch_bounds__(ReadsCompact2Cfg<W>::NT, 2) k_reads_scatter_compact2(ReadsSrc src, LevelArgs a, unsigned long long *__restrict__ cursor,
                                                                                     uint32_t *__restrict__ out, int64_t pos0, int64_t pos1) {
  extern __shared__ __align__(16) uint32_t smem[];
  using C = ReadsCompact2Cfg<W>;
  constexpr int NT = C::NT, T = C::T, SUB = C::SUB, NSUB = C::NSUB, KCAP = C::KCAP;
  static_assert(KCAP % NT == 0 && C::PCAP % 2 == 0 && T <= 65536, "compact2 tiling");
  const int nbins = 1 << a.nbits, tid = threadIdx.x, K1 = src.k + 1;
  uint32_t *s_cnt = smem;
  uint32_t *s_gd32 = s_cnt + nbins + 32;
  if ((reinterpret_cast<uintptr_t>(s_gd32) & 7) != 0) s_gd32 += 1;
  long long *s_gd = reinterpret_cast<long long *>(s_gd32);
  uint32_t *scratch = s_gd32 + 2 * nbins;
  uint32_t *s_pc = scratch + 36;
  uint16_t *s_pos = reinterpret_cast<uint16_t *>(s_pc + NT);
  uint32_t *stage = s_pc + NT + C::PCAP / 2;
  stage += ((16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15) >> 2;   // 16-byte aligned (staged_copy_out)
  uint32_t *seq = stage + (size_t)KCAP * W;
  uint32_t *sb = seq + reads_compact2_seq_words<W>();
  const int dsh = 32 - a.nbits;
  const uint32_t dlo = a.dlo, dspan = a.dhi - a.dlo;
  const bool ranged = dlo != 0u || a.dhi != (1u << a.nbits);
  const uint32_t dummy = (uint32_t)nbins + (tid & 31);

  const int64_t base = pos0 + (int64_t)blockIdx.x * T;
  {
    const int64_t total_words = (src.n_bases + 15) >> 4, w0 = base >> 4;
    const int nsw = reads_compact2_seq_words<W>();
    for (int i = tid; i < nsw; i += NT) seq[i] = (w0 + i < total_words) ? src.packed[w0 + i] : 0u;
    const int64_t total_bw = (src.n_bases + 31) >> 5, b0 = base >> 5;
    const int nbw = reads_compact2_bit_words(T, src.k);
    for (int i = tid; i < nbw; i += NT) sb[i] = (b0 + i < total_bw) ? src.sbits[b0 + i] : 0u;
  }
  int64_t lim64 = src.n_bases - K1 + 1 - base;
  if (pos1 - base < lim64) lim64 = pos1 - base;
  const int lim = (int)(lim64 < 0 ? 0 : (lim64 > T ? T : lim64));
  __syncthreads();

  // keys of the listed positions s_pos[j0 .. j0 + nb): counted, global ranges reserved, staged in bin order, copied out
  auto batch = [&](uint32_t j0, uint32_t nb) {
    for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
    __syncthreads();
    for (uint32_t j = tid; j < nb; j += NT) {
      const uint32_t d = reads_head_at(seq, s_pos[j0 + j], K1) >> dsh;
      const bool ok = !ranged || (d - dlo) < dspan;
      atomicAdd(s_cnt + (ok ? d : dummy), 1u);
    }
    __syncthreads();
    const uint32_t total = bins_scan_reserve<NT, 2>(s_cnt, s_gd, scratch, cursor, nbins, a.limit);
    for (uint32_t j = tid; j < nb; j += NT) {
      uint32_t key[W];
      reads_key_at<W>(seq, s_pos[j0 + j], K1, key);
      const uint32_t d = key[0] >> dsh;
      const bool ok = !ranged || (d - dlo) < dspan;
      const uint32_t pos = atomicAdd(s_cnt + (ok ? d : dummy), 1u);
      if (ok) {
        stage_store<W>(stage + (size_t)pos * W, key);
      }
    }
    __syncthreads();
    staged_copy_out<W, NT>(stage, s_gd, total, a, out, [dsh](const uint32_t *r) { return r[0] >> dsh; });
    __syncthreads();
  };

  uint32_t cnt = 0;   // listed and not yet scattered (block-uniform)
  for (int s = 0; s < NSUB; ++s) {
    if (s * SUB >= lim) break;
    const int p0 = s * SUB + tid * 16;
    const uint32_t vm = valid16(sb, p0, src.k, lim);
    s_pc[tid] = __popc(vm);
    __syncthreads();
    const uint32_t V = block_excl_scan<NT>(s_pc, NT, scratch);
    {
      uint32_t at = cnt + s_pc[tid];
      for (uint32_t m = vm; m; m &= m - 1) s_pos[at++] = (uint16_t)(p0 + __ffs(m) - 1);
    }
    cnt += V;
    __syncthreads();
    const bool last = s == NSUB - 1 || (s + 1) * SUB >= lim;
    uint32_t j0 = 0;
    while (cnt - j0 >= (uint32_t)KCAP || (last && cnt > j0)) {
      const uint32_t nb = min((uint32_t)KCAP, cnt - j0);
      batch(j0, nb);
      j0 += nb;
    }
    if (last) break;
    if (j0) {   // what the batches left over (< KCAP positions) moves to the front
      const uint32_t rest = cnt - j0;
      uint16_t tmp[KCAP / NT];
#pragma unroll
      for (int i = 0; i < KCAP / NT; ++i) {
        const uint32_t idx = (uint32_t)(i * NT + tid);
        tmp[i] = idx < rest ? s_pos[j0 + idx] : (uint16_t)0;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < KCAP / NT; ++i) {
        const uint32_t idx = (uint32_t)(i * NT + tid);
        if (idx < rest) s_pos[idx] = tmp[i];
      }
      cnt = rest;
      __syncthreads();
    }
  }
}

}  // namespace mf
