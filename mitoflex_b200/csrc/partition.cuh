// partition.cuh -- one MSD radix-partition level over W-word records.
//
// The same three kernels serve every level of the "bucketed" sort:
//   k_level_hist    : digit histogram per output segment          (read only)
//   k_level_scan    : histogram -> absolute cursors + bucket table (tiny)
//   k_level_scatter : rank in shared memory, stage, coalesced copy-out
// A Producer feeds a tile of records: ReadsProducer computes canonical (k+1)-mer keys straight from
// 2-bit packed reads (so unsorted keys are never written to HBM), RecordsProducer re-reads records that
// an earlier level wrote.  Ranking is one shared-memory atomicAdd per record on a CTA-wide counter array (keys only, so
// the partition need not be stable; on B200 that is ~20x cheaper than match.any, see tools/ubench.cu).
#pragma once
#include "common.cuh"

namespace mf {

// ---- chunk table: input chunks -> output segments (several chunks may feed one segment, which is how
// data received from several GPUs for the same prefix range is merged without an extra pass) ----------
struct ChunkTable {
  const int64_t *start;      // [nchunk] first record of the chunk in the input buffer
  const int64_t *size;       // [nchunk]
  const int32_t *seg;        // [nchunk] output segment id
  const int64_t *tile_base;  // [nchunk+1] exclusive prefix of ceil(size / T)
  int nchunk;
};

struct LevelArgs {
  int bit_off;          // bits already consumed above this digit
  int nbits;            // digit width, 1..kMaxDigitBits
  uint32_t dlo, dhi;    // keep only digits in [dlo, dhi) (out-of-core rounds / multi-GPU ownership)
  // Fused partition + exchange: when set, bin b is written at byte address bin_base[b] (+ cursor[b] records) instead of
  // into `out` -- the address may be a peer GPU's buffer mapped over NVLink, so keys go straight to their owner.
  const unsigned long long *bin_base;
  // Range partition (count level 2): segment s is cut into seg_nb[s] <= 2^nbits bins of equal key range instead of 2^nbits
  // bit-prefix bins, digit = (next xbits bits * seg_nb[s]) >> xbits -- monotone in the key, so bins stay key-ordered, and
  // every segment gets the bin count its size asks for (canonical keys are twice as dense at small prefixes).
  const uint16_t *seg_nb;
  int xbits;
};

template <int W>
__device__ __forceinline__ uint32_t level_digit(const uint32_t (&r)[W], const LevelArgs &a, uint32_t nb) {
  if (nb == 0u) return rec_digit<W>(r, a.bit_off, a.nbits);
  return (rec_digit<W>(r, a.bit_off, a.xbits) * nb) >> a.xbits;
}
template <int W>
__device__ __forceinline__ uint32_t level_digit_mem(const uint32_t *r, const LevelArgs &a, uint32_t nb) {
  if (nb == 0u) return rec_digit_mem<W>(r, a.bit_off, a.nbits);
  return (rec_digit_mem<W>(r, a.bit_off, a.xbits) * nb) >> a.xbits;
}

// one tile of a level launch: where its records are and which segment they belong to (built by k_build_tiles so that
// no CTA has to walk the chunk table with a chain of dependent loads)
struct TileDesc {
  int64_t base;
  int32_t n;
  int32_t seg;
};
__global__ void k_build_tiles(ChunkTable ct, int T, int64_t ntiles, TileDesc *out) {
  const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= ntiles) return;
  int lo = 0, hi = ct.nchunk;  // last chunk with tile_base <= tile
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (ct.tile_base[mid] <= tile) lo = mid; else hi = mid;
  }
  const int64_t off = (tile - ct.tile_base[lo]) * (int64_t)T;
  const int64_t rem = ct.size[lo] - off;
  TileDesc d;
  d.base = ct.start[lo] + off;
  d.n = (int32_t)(rem < T ? rem : T);
  d.seg = ct.seg[lo];
  out[tile] = d;
}

// ============================================================ producers
template <int W>
struct RecordsProducer {
  const uint32_t *in;
  const TileDesc *tiles;
  int T;
  static constexpr bool kNeedsSmem = false;
  __host__ __device__ static int smem_words(int, int) { return 4; }

  struct Tile {
    int64_t base;
    int n;
    int seg;
  };
  __device__ __forceinline__ Tile setup(int64_t tile, uint32_t *) const {
    const TileDesc d = tiles[tile];   // same address for the whole CTA: one broadcast load
    Tile t;
    t.base = d.base;
    t.n = d.n;
    t.seg = d.seg;
    __syncthreads();                  // the callers zero shared counters before setup()
    return t;
  }
  __device__ __forceinline__ bool get(const Tile &t, const uint32_t *, int j, uint32_t (&r)[W]) const {
    if (j >= t.n) return false;
    const uint32_t *p = in + (t.base + j) * (int64_t)W;
    if constexpr (W == 2) {
      uint2 v = *reinterpret_cast<const uint2 *>(p);
      r[0] = v.x; r[1] = v.y;
    } else if constexpr (W == 4) {
      uint4 v = *reinterpret_cast<const uint4 *>(p);
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) r[i] = p[i];
    }
    return true;
  }
};

// Canonical (k+1)-mer keys of the REVERSED read, computed per base position from true-orientation packed
// reads: stored edge = reverse(e), its reverse complement = complement(e); key = min of the two
// (megahit KmerCounter reads the library with is_reverse=true; strand tie -> the edge itself).
template <int W>
struct ReadsProducer {
  const uint32_t *packed;   // 16 bases / word, first base in the top bits, reads back to back
  const uint32_t *sbits;    // bit g (LSB-first within word) set iff a read starts at base g
  int64_t n_bases;
  int k;
  int T;
  static constexpr bool kNeedsSmem = true;
  __host__ __device__ static int n_seq_words(int T) { return T / 16 + W + 2; }
  __host__ __device__ static int n_bit_words(int T, int k) { return T / 32 + (k + 31) / 32 + 2; }
  __host__ __device__ static int smem_words(int T, int k) { return n_seq_words(T) + n_bit_words(T, k); }

  struct Tile {
    int64_t base;  // first base position of the tile
    int n;
    int seg;
  };
  __device__ __forceinline__ Tile setup(int64_t tile, uint32_t *sm) const {
    Tile t;
    t.base = tile * (int64_t)T;
    int64_t rem = n_bases - t.base;
    t.n = (int)(rem < T ? rem : T);
    t.seg = 0;
    const int64_t total_words = (n_bases + 15) >> 4;
    const int64_t w0 = t.base >> 4;  // T is a multiple of 32 so tiles are word aligned
    const int nsw = n_seq_words(T);
    for (int i = threadIdx.x; i < nsw; i += blockDim.x) sm[i] = (w0 + i < total_words) ? packed[w0 + i] : 0u;
    const int64_t total_bw = (n_bases + 31) >> 5;
    const int64_t b0 = t.base >> 5;
    const int nbw = n_bit_words(T, k);
    uint32_t *sb = sm + nsw;
    for (int i = threadIdx.x; i < nbw; i += blockDim.x) sb[i] = (b0 + i < total_bw) ? sbits[b0 + i] : 0u;
    __syncthreads();
    return t;
  }
  __device__ __forceinline__ bool get(const Tile &t, const uint32_t *sm, int j, uint32_t (&key)[W]) const {
    const int K1 = k + 1;
    if constexpr (W <= 2) {
      // the whole (k+1)-mer fits 64 bits (k <= 31): one 64-bit window, one validity funnel shift
      const uint32_t *sb = sm + n_seq_words(T);
      const int bit = j + 1;
      uint32_t v = __funnelshift_r(sb[bit >> 5], sb[(bit >> 5) + 1], bit & 31);
      v &= (1u << k) - 1u;                                   // k <= 31
      const bool valid = j < t.n && (t.base + j + K1 <= n_bases) && v == 0;
      const int wi = j >> 4, sh = (j & 15) * 2;
      const uint32_t w0 = sm[wi], w1 = sm[wi + 1], w2 = sm[wi + 2];
      const unsigned long long raw0 =
          ((unsigned long long)__funnelshift_l(w1, w0, sh) << 32) | __funnelshift_l(w2, w1, sh);
      const int bits = 2 * K1;
      const unsigned long long mask = ~0ull << (64 - bits);
      const unsigned long long raw = raw0 & mask;
      const unsigned long long cmpl = ~raw0 & mask;
      unsigned long long rv = __brevll(raw);                 // reversed bit string, right aligned
      rv = ((rv & 0x5555555555555555ull) << 1) | ((rv >> 1) & 0x5555555555555555ull);
      rv <<= (64 - bits);
      const unsigned long long kk = rv < cmpl ? rv : cmpl;
      key[0] = (uint32_t)(kk >> 32);
      if constexpr (W == 2) key[1] = (uint32_t)kk;
      return valid;
    }
    bool valid = j < t.n && (t.base + j + K1 <= n_bases);
    // no read may start inside (j, j+k]
    const uint32_t *sb = sm + n_seq_words(T);
    {
      int bit = j + 1, left = k;
      while (left > 0) {
        int wi = bit >> 5, sh = bit & 31;
        uint32_t v = __funnelshift_r(sb[wi], sb[wi + 1], sh);
        if (left < 32) v &= (1u << left) - 1u;
        valid = valid && (v == 0);
        bit += 32;
        left -= 32;
      }
    }
    // raw = 2*K1 bits at base j, left aligned
    const int wi = j >> 4, sh = (j & 15) * 2;
    uint32_t raw[W];
#pragma unroll
    for (int i = 0; i < W; ++i) raw[i] = __funnelshift_l(sm[wi + i + 1], sm[wi + i], sh);
    const int pad = 32 * W - 2 * K1;  // 0..30
    raw[W - 1] &= 0xffffffffu << pad;
    // complement(e): same direction, bases 3-c
    uint32_t cmpl[W];
#pragma unroll
    for (int i = 0; i < W; ++i) cmpl[i] = ~raw[i];
    cmpl[W - 1] &= 0xffffffffu << pad;
    // reverse(e): reverse base order (bit reverse, then swap the two bits of every base back)
    uint32_t rr[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      uint32_t x = __brev(raw[W - 1 - i]);
      rr[i] = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    }
    uint32_t rev[W];
#pragma unroll
    for (int i = 0; i < W; ++i) rev[i] = __funnelshift_l(i + 1 < W ? rr[i + 1] : 0u, rr[i], pad);
    // min(rev, cmpl)
    int c = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      if (c == 0 && rev[i] != cmpl[i]) c = rev[i] < cmpl[i] ? -1 : 1;
    }
#pragma unroll
    for (int i = 0; i < W; ++i) key[i] = c <= 0 ? rev[i] : cmpl[i];
    return valid;
  }
};

// ============================================================ histogram
// grid.x = number of tiles; dynamic smem = CTA histogram (u32 [nbins]) + producer words.
// Shared-memory atomics are the ranking primitive: measured on B200 (tools/ubench.cu) a warp-wide shared atomicAdd on
// scattered counters costs ~3 SM-cycles, match.any ~63 and a 10-ballot loop ~36 (both saturate the ADU pipe).
template <class P, int W, int NT, int IPT>
__global__ void __launch_bounds__(NT) k_level_hist(P prod, LevelArgs a, unsigned long long *__restrict__ hist) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int nbins = 1 << a.nbits;
  uint32_t *s_hist = smem;                 // [nbins]
  uint32_t *psm = smem + nbins;
  const int tid = threadIdx.x;
  for (int i = tid; i < nbins; i += NT) s_hist[i] = 0;
  typename P::Tile t = prod.setup(blockIdx.x, psm);   // contains a __syncthreads
  const uint32_t nb = a.seg_nb ? a.seg_nb[t.seg] : 0u;
#pragma unroll 4
  for (int i = 0; i < IPT; ++i) {
    uint32_t r[W];
    int j = i * NT + tid;
    bool valid = prod.get(t, psm, j, r);
    uint32_t d = level_digit<W>(r, a, nb);
    if (valid && d >= a.dlo && d < a.dhi) atomicAdd(s_hist + d, 1u);
  }
  __syncthreads();
  unsigned long long *h = hist + (size_t)t.seg * nbins;
  for (int b = tid; b < nbins; b += NT) {
    uint32_t c = s_hist[b];
    if (c) atomicAdd(h + b, (unsigned long long)c);
  }
}

// ============================================================ scan: hist -> cursors + bucket table
// seg_total[s] = sum_d hist[s][d]
__global__ void k_seg_totals(const unsigned long long *hist, int nbins, int nseg, int64_t *seg_total) {
  int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= nseg) return;
  unsigned long long acc = 0;
  for (int b = threadIdx.x & 31; b < nbins; b += 32) acc += hist[(size_t)s * nbins + b];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) seg_total[s] = (int64_t)acc;
}
// single-block exclusive scan of int64 v[0..n) -> out[0..n], out[n] = total (+base)
__global__ void k_scan_i64(const int64_t *v, int64_t n, int64_t base, int64_t *out) {
  __shared__ int64_t s_part[1024];
  __shared__ int64_t s_run;
  const int tid = threadIdx.x;
  if (tid == 0) s_run = base;
  __syncthreads();
  const int64_t per = (n + blockDim.x - 1) / blockDim.x;
  const int64_t b = tid * per, e = b + per < n ? b + per : n;
  int64_t sum = 0;
  for (int64_t i = b; i < e; ++i) sum += v[i];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int64_t run = s_run;
    for (int i = 0; i < (int)blockDim.x; ++i) { int64_t x = s_part[i]; s_part[i] = run; run += x; }
    out[n] = run;
  }
  __syncthreads();
  int64_t run = s_part[tid];
  for (int64_t i = b; i < e; ++i) { int64_t x = v[i]; out[i] = run; run += x; }
}
// per segment: cursor[s][d] = seg_out_start[s] + exclusive prefix of hist[s][.]; bucket table likewise.
__global__ void k_level_scan(const unsigned long long *hist, int nbins, const int64_t *seg_out_start,
                             unsigned long long *cursor, int64_t *bkt_start, int64_t *bkt_size) {
  // one block of 1024 threads per segment; each thread owns kMaxBins/1024 consecutive bins
  constexpr int PER = kMaxBins / 1024;
  __shared__ unsigned long long s_w[32];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long v[PER], sum = 0;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int b = tid * PER + q;
    v[q] = b < nbins ? hist[(size_t)s * nbins + b] : 0ull;
    sum += v[q];
  }
  unsigned long long inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long x = s_w[lane], xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    s_w[lane] = xi - x;
  }
  __syncthreads();
  unsigned long long run = (unsigned long long)seg_out_start[s] + s_w[warp] + inc - sum;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int b = tid * PER + q;
    if (b < nbins) {
      const size_t idx = (size_t)s * nbins + b;
      cursor[idx] = run;
      bkt_start[idx] = (int64_t)run;
      bkt_size[idx] = (int64_t)v[q];
      run += v[q];
    }
  }
}

// Thread t owns bins [t*BPT, (t+1)*BPT).  In: s_cnt = per-bin record counts of the tile.  Out: s_cnt = exclusive starts
// (staging offsets); s_gd[b] = global record index reserved for the bin's first staged record minus that start (one global
// atomicAdd per non-empty bin, all of a thread's in flight together).  Returns the tile's record total.
template <int NT, int BPT>
__device__ __forceinline__ uint32_t bins_scan_reserve(uint32_t *s_cnt, long long *s_gd, uint32_t *scratch,
                                                      unsigned long long *cursor_row, int nbins) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t c[BPT], sum = 0;
#pragma unroll
  for (int q = 0; q < BPT; ++q) {
    const int b = tid * BPT + q;
    c[q] = b < nbins ? s_cnt[b] : 0u;
    sum += c[q];
  }
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = lane < NT / 32 ? scratch[lane] : 0u;
    uint32_t vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
      if (lane >= o) vi += t;
    }
    scratch[lane] = vi - v;
    if (lane == 31) scratch[32] = vi;
  }
  __syncthreads();
  uint32_t run = scratch[warp] + inc - sum;
  unsigned long long g[BPT];
#pragma unroll
  for (int q = 0; q < BPT; ++q) {
    const int b = tid * BPT + q;
    g[q] = c[q] ? atomicAdd(cursor_row + b, (unsigned long long)c[q]) : 0ull;
  }
#pragma unroll
  for (int q = 0; q < BPT; ++q) {
    const int b = tid * BPT + q;
    if (b < nbins) {
      s_cnt[b] = run;
      s_gd[b] = (long long)g[q] - (long long)run;
      run += c[q];
    }
  }
  const uint32_t total = scratch[32];
  __syncthreads();
  return total;
}

// ============================================================ scatter
// dynamic smem layout (uint32 units):
//   s_cnt   u32 [nbins]      per-bin counters (rank = atomicAdd), then exclusive starts
//   s_gd    i64 [nbins]      global record index of the bin's first staged record minus its staged index
//   scratch u32 [34]
//   stage   u32 [T*W]
//   producer words
template <int W>
__host__ __device__ inline size_t scatter_smem_bytes(int NT, int T, int nbits, int prod_words) {
  size_t nb = (size_t)1 << nbits;
  size_t words = nb + 2 * nb + 34 + 4 + (size_t)T * W + prod_words;
  return words * 4;
}

template <class P, int W, int NT, int IPT, int BPT>
__global__ void __launch_bounds__(NT) k_level_scatter(P prod, LevelArgs a, unsigned long long *__restrict__ cursor,
                                                      uint32_t *__restrict__ out) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int T = NT * IPT;
  const int nbins = 1 << a.nbits;
  const int tid = threadIdx.x;
  uint32_t *s_cnt = smem;
  uint32_t *s_gd32 = s_cnt + nbins;
  if ((reinterpret_cast<uintptr_t>(s_gd32) & 7) != 0) s_gd32 += 1;   // 8-byte align
  long long *s_gd = reinterpret_cast<long long *>(s_gd32);
  uint32_t *scratch = s_gd32 + 2 * nbins;
  uint32_t *stage = scratch + 34;
  if ((reinterpret_cast<uintptr_t>(stage) & 7) != 0) stage += 1;
  uint32_t *psm = stage + (size_t)T * W;

  for (int i = tid; i < nbins; i += NT) s_cnt[i] = 0;
  typename P::Tile t = prod.setup(blockIdx.x, psm);   // contains a __syncthreads

  const uint32_t nb = a.seg_nb ? a.seg_nb[t.seg] : 0u;
  uint32_t rec[IPT][W];
  uint32_t rk[IPT];   // digit << 16 | rank within the bin; 0xffffffff = dropped
  bool ok[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) ok[i] = prod.get(t, psm, i * NT + tid, rec[i]);   // all loads in flight before any atomic
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t d = level_digit<W>(rec[i], a, nb);
    const bool valid = ok[i] && d >= a.dlo && d < a.dhi;
    rk[i] = valid ? ((d << 16) | atomicAdd(s_cnt + d, 1u)) : 0xffffffffu;
  }
  __syncthreads();
  const uint32_t total = bins_scan_reserve<NT, BPT>(s_cnt, s_gd, scratch, cursor + (size_t)t.seg * nbins, nbins);
  // stage records in bin order
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    if (rk[i] != 0xffffffffu) {
      uint32_t pos = s_cnt[rk[i] >> 16] + (rk[i] & 0xffffu);
#pragma unroll
      for (int c = 0; c < W; ++c) stage[(size_t)pos * W + c] = rec[i][c];
    }
  }
  __syncthreads();
  // coalesced copy-out: consecutive staged records of a bin go to consecutive global records
  if constexpr (W == 2) {
    const uint2 *st2 = reinterpret_cast<const uint2 *>(stage);
    uint2 *out2 = reinterpret_cast<uint2 *>(out);
    for (uint32_t j = tid; j < total; j += NT) {
      uint2 v = st2[j];
      uint32_t r2[2] = {v.x, v.y};
      uint32_t d = level_digit<2>(r2, a, nb);
      uint2 *dst = a.bin_base ? reinterpret_cast<uint2 *>(a.bin_base[d]) : out2;
      dst[s_gd[d] + (long long)j] = v;
    }
  } else {
    const uint32_t total_words = total * W;
    for (uint32_t x = tid; x < total_words; x += NT) {
      uint32_t j = x / W, c = x - j * W;
      uint32_t d = level_digit_mem<W>(stage + (size_t)j * W, a, nb);
      uint32_t *dst = a.bin_base ? reinterpret_cast<uint32_t *>(a.bin_base[d]) : out;
      dst[(s_gd[d] + (long long)j) * W + c] = stage[x];
    }
  }
}

}  // namespace mf
