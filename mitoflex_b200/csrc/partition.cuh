// partition.cuh -- one MSD radix-partition level over W-word records.
//
// The same three kernels serve every level of the "bucketed" sort:
//   k_level_hist    : digit histogram per output segment          (read only)
//   k_level_scan    : histogram -> absolute cursors + bucket table (tiny)
//   k_level_scatter : rank in shared memory, stage, coalesced copy-out
// A Producer feeds a tile of records: RecordsProducer re-reads records that an earlier level wrote (the reads-fed level,
// which computes canonical (k+1)-mer keys straight from 2-bit packed reads, has its own kernels in reads.cuh).  Ranking is one shared-memory atomicAdd per record on a CTA-wide counter array (keys only, so
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
  // Sampled-histogram mode of the reads-fed level: bin b owns [its start, limit[b]) and a tile's run that would cross the
  // limit is dropped (the host sees cursor > limit afterwards and redoes the level with an exact histogram).
  const unsigned long long *limit;
};
constexpr long long kDropRun = (long long)0x8000000000000000ull;

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
static __global__ void k_build_tiles(ChunkTable ct, int T, int64_t ntiles, TileDesc *out) {
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


// Persistent variant for records: a CTA walks a contiguous range of tiles, keeps its shared histogram across the tiles of
// one segment (tiles are segment-major, so it flushes a handful of times instead of once per tile) and has all the loads
// of a round in flight before the first shared atomic.  The per-tile kernel above waits on a chain of dependent loads
// (tile descriptor -> records) at the start of every short-lived CTA (long_scoreboard 73 % under ncu).
template <int W, int NT>
__global__ void __launch_bounds__(NT) k_level_hist_persist(const uint32_t *__restrict__ in, const TileDesc *__restrict__ tiles, int64_t ntiles,
                                                           LevelArgs a, unsigned long long *__restrict__ hist) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int IPT = W <= 2 ? 16 : (W <= 4 ? 8 : 4);   // records held in registers per round
  const int nbins = 1 << a.nbits;
  uint32_t *s_hist = smem;
  const int tid = threadIdx.x;
  const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
  if (t0 >= t1) return;
  for (int i = tid; i < nbins; i += NT) s_hist[i] = 0;
  int cur_seg = -1;
  uint32_t nb = 0;
  TileDesc d = tiles[t0];
  auto flush = [&]() {
    __syncthreads();
    if (cur_seg >= 0) {
      unsigned long long *h = hist + (size_t)cur_seg * nbins;
      for (int b = tid; b < nbins; b += NT) {
        const uint32_t c = s_hist[b];
        if (c) {
          atomicAdd(h + b, (unsigned long long)c);
          s_hist[b] = 0;
        }
      }
    }
    __syncthreads();
  };
  for (int64_t t = t0; t < t1; ++t) {
    const TileDesc nd = t + 1 < t1 ? tiles[t + 1] : d;   // the next descriptor is on its way while this tile is read
    if (d.seg != cur_seg) {
      flush();
      cur_seg = d.seg;
      nb = a.seg_nb ? a.seg_nb[cur_seg] : 0u;
    }
    if constexpr (W == 2) {
      // 16-byte loads: two records per lane; the tile is read from the 16-byte boundary at or below its first record.
      // Lean digit (bit_off < 32, checked by the host): (x * nb) >> xbits, bit-prefix levels pass nb = 2^xbits.
      const int off = (int)(d.base & 1);
      const uint4 *p4 = reinterpret_cast<const uint4 *>(in + (d.base - off) * 2);
      const int npair = (d.n + off + 1) >> 1;
      const int xb = nb ? a.xbits : a.nbits, shx = 32 - xb;
      const uint32_t mul = nb ? nb : (1u << a.nbits);
      const bool whole = off == 0 && (d.n & 1) == 0 && npair % (NT * 8) == 0;
      for (int e0 = 0; e0 < npair; e0 += NT * 8) {
        uint4 v[8];
        if (whole) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = p4[e0 + i * NT + tid];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            atomicAdd(s_hist + (((__funnelshift_l(v[i].y, v[i].x, a.bit_off) >> shx) * mul) >> xb), 1u);
            atomicAdd(s_hist + (((__funnelshift_l(v[i].w, v[i].z, a.bit_off) >> shx) * mul) >> xb), 1u);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = p4[min(e0 + i * NT + tid, npair - 1)];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int e = e0 + i * NT + tid;
            const int j = 2 * e - off;   // record index of v[i].xy within the tile; .zw is j + 1
            const uint32_t d0 = ((__funnelshift_l(v[i].y, v[i].x, a.bit_off) >> shx) * mul) >> xb;
            const uint32_t d1 = ((__funnelshift_l(v[i].w, v[i].z, a.bit_off) >> shx) * mul) >> xb;
            if (e < npair && j >= 0) atomicAdd(s_hist + d0, 1u);
            if (e < npair && j + 1 < d.n) atomicAdd(s_hist + d1, 1u);
          }
        }
      }
    } else {
    const uint32_t *base = in + d.base * (int64_t)W;
    for (int j0 = 0; j0 < d.n; j0 += NT * IPT) {
      uint32_t r[IPT][W];
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        const int j = min(j0 + i * NT + tid, d.n - 1);   // clamped: a duplicate load instead of a branch, dropped below
        if constexpr (W == 2) {
          const uint2 v = *reinterpret_cast<const uint2 *>(base + (size_t)j * 2);
          r[i][0] = v.x; r[i][1] = v.y;
        } else if constexpr (W == 4) {
          const uint4 v = *reinterpret_cast<const uint4 *>(base + (size_t)j * 4);
          r[i][0] = v.x; r[i][1] = v.y; r[i][2] = v.z; r[i][3] = v.w;
        } else {
#pragma unroll
          for (int c = 0; c < W; ++c) r[i][c] = base[(size_t)j * W + c];
        }
      }
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        const uint32_t dg = level_digit<W>(r[i], a, nb);
        if (j0 + i * NT + tid < d.n && dg >= a.dlo && dg < a.dhi) atomicAdd(s_hist + dg, 1u);
      }
    }
    }
    d = nd;
  }
  flush();
}

// ============================================================ scan: hist -> cursors + bucket table
// seg_total[s] = sum_d hist[s][d]
static __global__ void k_seg_totals(const unsigned long long *hist, int nbins, int nseg, int64_t *seg_total) {
  int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= nseg) return;
  unsigned long long acc = 0;
  for (int b = threadIdx.x & 31; b < nbins; b += 32) acc += hist[(size_t)s * nbins + b];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) seg_total[s] = (int64_t)acc;
}
// single-block exclusive scan of int64 v[0..n) -> out[0..n], out[n] = total (+base)
static __global__ void k_scan_i64(const int64_t *v, int64_t n, int64_t base, int64_t *out) {
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
static __global__ void k_level_scan(const unsigned long long *hist, int nbins, const int64_t *seg_out_start,
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
                                                      unsigned long long *cursor_row, int nbins,
                                                      const unsigned long long *limit = nullptr) {
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
      s_gd[b] = (limit && c[q] && g[q] + c[q] > limit[b]) ? kDropRun : (long long)g[q] - (long long)run;
      run += c[q];
    }
  }
  const uint32_t total = scratch[32];
  __syncthreads();
  return total;
}

// a W-word record into its staging slot (16-byte aligned staging area): 16- / 8-byte stores where the record size allows.  Word
// stores of 8-word records hit 4 of the 32 banks per instruction (slot * 8 + c): an 8-way conflict on every store, and
// `short_scoreboard` / `mio_throttle` on top of the stall list of the wide-key scatters (ncu r2ai); records of an odd number of
// words spread over all banks by themselves.
template <int W>
__device__ __forceinline__ void stage_store(uint32_t *dst, const uint32_t (&rec)[W]) {
  if constexpr (W >= 3 && W % 4 == 0) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) reinterpret_cast<uint4 *>(dst)[q] = make_uint4(rec[4 * q], rec[4 * q + 1], rec[4 * q + 2], rec[4 * q + 3]);
  } else if constexpr (W >= 3 && W % 2 == 0) {
#pragma unroll
    for (int q = 0; q < W / 2; ++q) reinterpret_cast<uint2 *>(dst)[q] = make_uint2(rec[2 * q], rec[2 * q + 1]);
  } else {
#pragma unroll
    for (int c = 0; c < W; ++c) dst[c] = rec[c];
  }
}

// ============================================================ copy-out of staged records
// `total` records of W words lie in `stage` in bin order; bin d's run goes to global record s_gd[d] + (staged index).  A record is
// moved by W / VEC threads in chunks of VEC words (4 when the record size is a multiple of 16 bytes, 2 of 8, else 1): the threads
// of a warp still write consecutive words, the bin and its global offset are looked up once per chunk instead of once per word
// and there is no division in the loop (a word-per-thread loop spent 30 instructions per word: a third of the compacted
// reads-fed scatter at k = 119, ncu r2ab).  `stage` must be 16-byte aligned; destinations that are not take words.
template <int W, int NT, class DigitOf>
__device__ __forceinline__ void staged_copy_out(const uint32_t *stage, const long long *s_gd, uint32_t total, const LevelArgs &a,
                                                uint32_t *__restrict__ out, DigitOf digit_of) {
  constexpr int VEC = (W % 4 == 0) ? 4 : ((W % 2 == 0) ? 2 : 1);
  constexpr int CPK = W / VEC;    // chunks per record
  constexpr int G = NT / CPK;     // records per sweep of the block
  const uint32_t tid = threadIdx.x;
  if (tid >= (uint32_t)(G * CPK)) return;
  const uint32_t g = tid / CPK, c = tid - g * CPK;
  for (uint32_t j = g; j < total; j += G) {
    const uint32_t *src = stage + (size_t)j * W;
    const uint32_t d = digit_of(src);
    const long long gd = s_gd[d];
    if (gd == kDropRun) continue;
    uint32_t *dst = (a.bin_base ? reinterpret_cast<uint32_t *>(a.bin_base[d]) : out) + (gd + (long long)j) * W + c * VEC;
    if constexpr (VEC == 4) {
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src + c * 4);
        continue;
      }
    } else if constexpr (VEC == 2) {
      if ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
        *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(src + c * 2);
        continue;
      }
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) dst[q] = src[c * VEC + q];
  }
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
  stage += ((16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15) >> 2;   // 16-byte aligned (staged_copy_out)
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
      stage_store<W>(stage + (size_t)pos * W, rec[i]);
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
    staged_copy_out<W, NT>(stage, s_gd, total, a, out, [&](const uint32_t *r) { return level_digit_mem<W>(r, a, nb); });
  }
}

}  // namespace mf
