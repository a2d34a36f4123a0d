// skm.cu -- super-k-mer exchange for the multi-GPU count (16 <= k <= 26, i.e. 2-word keys that leave room for a run).
//
// The fused partition + exchange of reads.cuh sends ONE 8-byte key per (k+1)-mer over NVLink: 29 GB out of every GPU at 8
// GPUs, and the exchange is link-bound (SCALE_r01: 50 ms of a 190 ms step).  Here the owner of a (k+1)-mer is a hash of its
// MINIMIZER (the smallest hashed strand-invariant m-mer inside it), so consecutive (k+1)-mers of a read share their owner
// most of the time, and a whole run travels as ONE 64-bit record: the run's bases (2 bits each, <= 30) plus its length.
// ~5 (k+1)-mers per record at k=21: a fifth of the bytes.  Every GPU then owns a pseudo-random, well balanced subset of the
// canonical keys (no +20 % on rank 0 from the density of canonical keys at small prefixes), receives records from every
// source in its own region of the owner's buffer, and runs the single-GPU count on them: the level-1 partition below
// expands the records into canonical keys (exactly KeyWindow's keys, then mixed by one Feistel round -- see skm_mix_hi),
// everything after it is count_finish unchanged.  The ranks' edge sets are disjoint, interleaved in key order and each in
// mixed-key order; the sdbg stage routes items by their own prefix, so it does not care.
//
//   k_skm_scatter  : reads -> minimizer owner per position -> runs -> records staged by owner -> copied out to the owners
//   k_skm_hist     : (sampled) level-1 digit histogram of the keys inside received records
//   k_skm_kscatter : records -> canonical keys scattered by their level-1 digit (the stand-in for k_reads_scatter)
//
// megahit has no counterpart (one process, shared memory): this replaces the all-to-all of SURVEY 8e.
#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include "engine.cuh"
#include "partition.cuh"
#include "reads.cuh"

namespace mf {

constexpr int kSkmNT = 512;
constexpr int kSkmMaxDst = 16;
constexpr int kSkmTileKeys = kSkmNT * 16;   // key slots of a receiver tile: 16 per thread = 16 / SLOTS records of <= SLOTS keys

// (k+1)-mers a record may hold: its bases must fit 60 bits, its length 3.  MFSDBG_SKM_CMAX (2..8) lowers it: shorter records
// cost NVLink bytes but fill the receiver's per-record key slots better (a record of <= 4 keys takes the 4-slot kernel).
static int skm_cmax(int K1) {
  int c = 31 - K1 > 8 ? 8 : 31 - K1;
  const char *e = getenv("MFSDBG_SKM_CMAX");
  if (e && *e) c = std::max(2, std::min(c, atoi(e)));
  return c;
}
static int skm_slots(int K1) { return skm_cmax(K1) <= 4 ? 4 : 8; }
bool skm_supported(int k) { return k >= 16 && k <= 26; }
int64_t skm_key_capacity(int64_t n_keys) { return (int64_t)(1.12 * (double)n_keys) + (int64_t)131072 * 1024 + 64; }

template <class K>
static void skm_set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) throw std::runtime_error("kernel shared memory request exceeds 227 KB");
  MF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

struct SkmArgs {
  int n_dst, m, cmax;
  unsigned long long dst[kSkmMaxDst];   // byte address of this source's region in destination d's buffer
  unsigned long long cap[kSkmMaxDst];   // records the region holds
  unsigned long long *cursor;           // [2 * n_dst]: records written so far, then keys, per destination
};

// owner (4 bits each) of the 16 (k+1)-mers that start in the thread's packed word: the minimizer of position i is the smallest
// hash among the NM m-mers at offsets i .. i + NM - 1.  An m-mer enters as min(complement, reverse) of its 2m bits -- the same
// strand symmetry as the canonical key, so a (k+1)-mer and its reverse complement agree on the minimizer, hence on the owner.
template <int NM>
__device__ __forceinline__ unsigned long long skm_owners(const uint32_t *w, int m, int n_dst, uint32_t &chg) {
  constexpr int NH = 16 + NM - 1;
  const uint32_t cw0 = ~w[0], cw1 = ~w[1], cw2 = ~w[2];
  const uint32_t cwv[3] = {cw0, cw1, cw2};
  // reversed stream (base b at position 47 - b), moved left by 16 - m bases: the m-mer at offset j then starts at position 32 - j
  const uint32_t r0 = rev_bases(w[2]), r1 = rev_bases(w[1]), r2 = rev_bases(w[0]);
  const int q = 2 * (16 - m);
  const uint32_t rp[4] = {__funnelshift_l(r1, r0, q), __funnelshift_l(r2, r1, q), __funnelshift_l(0u, r2, q), 0u};
  const int dn = 32 - 2 * m;
  uint32_t h[NH];
#pragma unroll
  for (int j = 0; j < NH; ++j) {
    const uint32_t cj = __funnelshift_l(cwv[(j >> 4) + 1], cwv[j >> 4], 2 * (j & 15)) >> dn;
    const int p = 32 - j;
    const uint32_t rj = __funnelshift_l(rp[(p >> 4) + 1], rp[p >> 4], 2 * (p & 15)) >> dn;
    h[j] = min(cj, rj) * 0x9E3779B1u;
  }
  // sliding minimum over windows of NM (van Herk): pre[j] = min of h from its block start to j, suf[j] = min from j to its block end
  uint32_t pre[NH], suf[NH];
#pragma unroll
  for (int j = 0; j < NH; ++j) pre[j] = (j % NM == 0) ? h[j] : min(pre[j - 1], h[j]);
#pragma unroll
  for (int j = NH - 1; j >= 0; --j) suf[j] = (j % NM == NM - 1 || j == NH - 1) ? h[j] : min(suf[j + 1], h[j]);
  // The owner comes from the LOW 20 bits of the winning hash: taking the minimum biases the top bits towards zero and leaves
  // the low ones uniform.  chg: bit i set when position i's owner differs from position i - 1's.
  unsigned long long own = 0;
  uint32_t prev = 0xffffffffu, ch = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t o = __umulhi(min(suf[i], pre[i + NM - 1]) << 12, (uint32_t)n_dst);
    own |= (unsigned long long)o << (4 * i);
    ch |= (o != prev ? 1u : 0u) << i;
    prev = o;
  }
  chg = ch;
  return own;
}

// dynamic smem (uint32 units): s_cnt[256] (owner-major: owner * 16 + warp) | s_keys[256] | s_start[17] | pad | s_gd i64[16] |
// scratch[36] | stage u64[T] | seq | sb
__host__ __device__ inline size_t skm_scatter_smem_bytes(int k, bool count_only) {
  const size_t T = (size_t)kSkmNT * 16;
  return (256 + 256 + 18 + 2 + 32 + 36 + 2 + (count_only ? 0 : 2 * T) + reads_seq_words(kSkmNT, 2) + reads_bit_words(kSkmNT, k)) * 4;
}

template <int NM, bool COUNT_ONLY>
__global__ void __launch_bounds__(kSkmNT, 2) k_skm_scatter(ReadsSrc src, SkmArgs a, int64_t ntiles, int64_t tstride) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NT = kSkmNT, T = NT * 16;
  uint32_t *s_cnt = smem;            // [16 owners][16 warps]
  uint32_t *s_keys = s_cnt + 256;
  uint32_t *s_start = s_keys + 256;  // [17]
  long long *s_gd = reinterpret_cast<long long *>(s_start + 18 + ((reinterpret_cast<uintptr_t>(s_start + 18) & 7) ? 1 : 0));
  uint32_t *scratch = reinterpret_cast<uint32_t *>(s_gd + 16);
  uint32_t *stage32 = scratch + 36;
  if ((reinterpret_cast<uintptr_t>(stage32) & 7) != 0) stage32 += 1;
  unsigned long long *stage = reinterpret_cast<unsigned long long *>(stage32);
  uint32_t *seq = stage32 + (COUNT_ONLY ? 0 : 2 * T);
  uint32_t *sb = seq + reads_seq_words(NT, 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K1 = src.k + 1;

  for (int64_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const int64_t tile = ti * tstride;
    __syncthreads();   // the previous tile's copy-out has read the stage and the tables
    const int lim = reads_load_tile<2, NT>(src, tile, seq, sb);
    s_cnt[tid & 255] = 0;
    s_keys[tid & 255] = 0;   // (both halves of the CTA write the same zeros)
    __syncthreads();
    const uint32_t vm = valid16(sb, tid * 16, src.k, lim);
    unsigned long long own = 0;
    uint32_t bm = 0;   // bit i: a record starts at position i
    if (vm) {
      uint32_t chg;
      own = skm_owners<NM>(seq + tid, a.m, a.n_dst, chg);
      // a record starts where a run of valid positions with one owner starts, and every cmax positions inside a run
      const uint32_t ns = vm & (~(vm << 1) | chg);   // natural starts
      const uint32_t cont = vm & ~ns;                // positions that continue a run
      uint32_t all = cont;                           // bit p: cont at p, p - 1, ..., p - cmax + 1
#pragma unroll
      for (int t = 1; t < 8; ++t)
        if (t < a.cmax) all &= cont << t;
      bm = ns;
      for (uint32_t f = (ns << a.cmax) & all; f & 0xffffu; f = (f << a.cmax) & all) bm |= f;
      bm &= 0xffffu;
    }
    const uint32_t stop = bm | (~vm & 0xffffu) | 0x10000u;   // where a run ends: the next record, an invalid position, the word's end
    // phase A: records and keys per (owner, warp)
    for (uint32_t rest = bm; rest;) {
      const int i = __ffs(rest) - 1;
      rest &= rest - 1;
      const uint32_t o = (uint32_t)(own >> (4 * i)) & 15u;
      const uint32_t cnt = (uint32_t)__ffs(stop >> (i + 1));
      atomicAdd(s_cnt + o * 16 + warp, 1u);
      atomicAdd(s_keys + o * 16 + warp, cnt);
    }
    __syncthreads();
    uint32_t ksum = 0;
    if (tid < a.n_dst) {
#pragma unroll
      for (int q = 0; q < 16; ++q) ksum += s_keys[tid * 16 + q];
    }
    const uint32_t total = block_excl_scan<NT>(s_cnt, 256, scratch);   // staging offset of every (owner, warp)
    if (tid < 16) {
      const uint32_t st = s_cnt[tid * 16], en = tid == 15 ? total : s_cnt[(tid + 1) * 16], n = en - st;
      s_start[tid] = st;
      if (tid == 15) s_start[16] = total;
      long long gd = kDropRun;
      if (tid < a.n_dst && n) {
        const unsigned long long g = atomicAdd(a.cursor + tid, (unsigned long long)n);
        if (!COUNT_ONLY && g + n <= a.cap[tid]) gd = (long long)g - (long long)st;
      }
      if (tid < a.n_dst && ksum) atomicAdd(a.cursor + a.n_dst + tid, (unsigned long long)ksum);
      s_gd[tid] = gd;
    }
    if constexpr (!COUNT_ONLY) {
      __syncthreads();
      // phase B: the records, staged by owner
      if (bm) {
        const uint32_t w0 = seq[tid], w1 = seq[tid + 1], w2 = seq[tid + 2];
        for (uint32_t rest = bm; rest;) {
          const int i = __ffs(rest) - 1;
          rest &= rest - 1;
          const uint32_t o = (uint32_t)(own >> (4 * i)) & 15u;
          const uint32_t cnt = (uint32_t)__ffs(stop >> (i + 1));
          const uint32_t pos = atomicAdd(s_cnt + o * 16 + warp, 1u);
          const uint32_t hi = __funnelshift_l(w1, w0, 2 * i), lo = __funnelshift_l(w2, w1, 2 * i);
          const int L = K1 + (int)cnt - 1;
          const unsigned long long bases = (((unsigned long long)hi << 32) | lo) & (~0ull << (64 - 2 * L));
          stage[pos] = bases | (unsigned long long)(cnt - 1);
        }
      }
      __syncthreads();
      // copy-out: the warps share the owners' runs; consecutive staged records go to consecutive records of the region
      {
        const int o = warp % a.n_dst, part = warp / a.n_dst, nparts = (16 - o + a.n_dst - 1) / a.n_dst;
        const long long gd = s_gd[o];
        if (gd != kDropRun) {
          const uint32_t st = s_start[o], en = s_start[o + 1];
          unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.dst[o]);
          for (uint32_t j = st + (uint32_t)(part * 32 + lane); j < en; j += (uint32_t)(nparts * 32)) dst[gd + (long long)j] = stage[j];
        }
      }
    }
  }
}

// ---- receiver: canonical keys out of a record (bit-identical to KeyWindow<2>::key at the same positions)
__device__ __forceinline__ unsigned long long skm_rev64(unsigned long long b) {
  return ((unsigned long long)rev_bases((uint32_t)b) << 32) | rev_bases((uint32_t)(b >> 32));
}
// One Feistel round over the 2*K1 key bits (left-aligned in hi:lo): the top K1 bits are XORed with a hash of the low K1 bits.
// A bijection and its own inverse.  Why: a GPU owns the (k+1)-mers of certain minimizers, and an m-mer with a small hash is
// the minimizer of nearly every (k+1)-mer that STARTS with it -- so in key order a rank's keys pile up 8-fold (at 8 GPUs)
// inside some prefix ranges and are missing from others, and the equal-key-range buckets of the count overflow their tables
// (460 ms in the fallback sort at 8 GPUs, r2s).  Counting the mixed keys instead makes every bucket see the average
// density, whatever the minimizers (and the 2x density of canonical keys at small prefixes goes away as well); the edges
// come out in mixed order and are mapped back once (k_skm_unmix_edges) -- the sdbg stage routes their items by prefix anyway.
__device__ __forceinline__ uint32_t skm_mix_hi(uint32_t hi, uint32_t lo, int K1) {
  const uint32_t r = __funnelshift_l(lo, hi, K1) >> (32 - K1);
  return hi ^ (((r * 0x9E3779B1u) >> (32 - K1)) << (32 - K1));
}
struct SkmKeys {
  unsigned long long cb, rv, mask;
  int rsh, K1;   // rsh = 2 * (32 - K1)
  __device__ __forceinline__ void init(unsigned long long rec, int K1_) {
    const unsigned long long b = rec & ~7ull;
    K1 = K1_;
    cb = ~b;
    rv = skm_rev64(b);
    mask = ~0ull << (64 - 2 * K1);
    rsh = 2 * (32 - K1);
  }
  // canonical key of the record's i-th (k+1)-mer (bit-identical to KeyWindow<2>::key), mixed
  __device__ __forceinline__ unsigned long long key(int i) const {
    const unsigned long long c = (cb << (2 * i)) & mask, r = (rv << (rsh - 2 * i)) & mask;
    const unsigned long long v = c < r ? c : r;
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    return ((unsigned long long)skm_mix_hi(hi, lo, K1) << 32) | lo;
  }
};
// edge records (key bits on top of words 0-1, multiplicity in the low 16 bits of the last word) back to true keys
__global__ void k_skm_unmix_edges(uint32_t *edges, int64_t n, int We, int K1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t *e = edges + i * We;
  e[0] = skm_mix_hi(e[0], e[1], K1);
}

// records of tile `d`, R per thread; a record beyond the tile reads as "no keys"
template <int R>
__device__ __forceinline__ void skm_load(const unsigned long long *recs, const TileDesc &d, int tid, unsigned long long (&rec)[R],
                                         uint32_t (&cnt)[R]) {
#pragma unroll
  for (int q = 0; q < R; ++q) {
    const int j = q * kSkmNT + tid;
    rec[q] = j < d.n ? recs[d.base + j] : 0ull;
    cnt[q] = j < d.n ? (uint32_t)(rec[q] & 7ull) + 1u : 0u;
  }
}

template <int SLOTS>
__global__ void __launch_bounds__(kSkmNT) k_skm_hist(const unsigned long long *__restrict__ recs, const TileDesc *__restrict__ tiles,
                                                     int64_t ntiles, int64_t tstride, int K1, int nbits,
                                                     unsigned long long *__restrict__ hist) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int nbins = 1 << nbits, tid = threadIdx.x;
  uint32_t *s_hist = smem;   // [nbins + 32]
  const uint32_t dummy = (uint32_t)nbins + (tid & 31);
  const int dsh = 64 - nbits;
  for (int i = tid; i < nbins + 32; i += kSkmNT) s_hist[i] = 0;
  __syncthreads();
  for (int64_t ti = blockIdx.x; ti * tstride < ntiles; ti += gridDim.x) {
    const TileDesc d = tiles[ti * tstride];
    constexpr int R = 16 / SLOTS;
    unsigned long long rec[R];
    uint32_t cnt[R];
    skm_load<R>(recs, d, tid, rec, cnt);
#pragma unroll
    for (int q = 0; q < R; ++q) {
      SkmKeys sk;
      sk.init(rec[q], K1);
#pragma unroll
      for (int i = 0; i < SLOTS; ++i) {
        const uint32_t dg = (uint32_t)(sk.key(i) >> dsh);
        atomicAdd(s_hist + ((uint32_t)i < cnt[q] ? dg : dummy), 1u);
      }
    }
  }
  __syncthreads();
  for (int b = tid; b < nbins; b += kSkmNT) {
    const uint32_t c = s_hist[b];
    if (c) atomicAdd(hist + b, (unsigned long long)c);
  }
}

// dynamic smem (uint32 units): s_cnt[nbins+32] | pad | s_gd i64[nbins] | scratch[36] | pad | stage[8192 * 2]
__host__ __device__ inline size_t skm_kscatter_smem_bytes(int nbits) {
  const size_t nb = (size_t)1 << nbits;
  return (nb + 32 + 2 + 2 * nb + 36 + 2 + (size_t)kSkmTileKeys * 2) * 4;
}
template <int BPT, int SLOTS>
__global__ void __launch_bounds__(kSkmNT, 2) k_skm_kscatter(const unsigned long long *__restrict__ recs, const TileDesc *__restrict__ tiles,
                                                         int K1, int nbits, unsigned long long *__restrict__ cursor,
                                                         const unsigned long long *__restrict__ limit, uint32_t *__restrict__ out) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NT = kSkmNT;
  const int nbins = 1 << nbits, tid = threadIdx.x;
  uint32_t *s_cnt = smem;
  uint32_t *s_gd32 = s_cnt + nbins + 32;
  if ((reinterpret_cast<uintptr_t>(s_gd32) & 7) != 0) s_gd32 += 1;
  long long *s_gd = reinterpret_cast<long long *>(s_gd32);
  uint32_t *scratch = s_gd32 + 2 * nbins;
  uint32_t *stage = scratch + 36;
  if ((reinterpret_cast<uintptr_t>(stage) & 7) != 0) stage += 1;
  const uint32_t dummy = (uint32_t)nbins + (tid & 31);
  const int dsh = 64 - nbits;
  for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
  const TileDesc d = tiles[blockIdx.x];
  constexpr int R = 16 / SLOTS;
  unsigned long long rec[R];
  uint32_t cnt[R];
  skm_load<R>(recs, d, tid, rec, cnt);
  SkmKeys sk[R];
#pragma unroll
  for (int q = 0; q < R; ++q) sk[q].init(rec[q], K1);
  __syncthreads();
  // phase A: count
#pragma unroll
  for (int q = 0; q < R; ++q) {
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
      const uint32_t dg = (uint32_t)(sk[q].key(i) >> dsh);
      atomicAdd(s_cnt + ((uint32_t)i < cnt[q] ? dg : dummy), 1u);
    }
  }
  __syncthreads();
  const uint32_t total = bins_scan_reserve<NT, BPT>(s_cnt, s_gd, scratch, cursor, nbins, limit);
  // phase B: keys again, each takes the next free slot of its bin
#pragma unroll
  for (int q = 0; q < R; ++q) {
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
      const unsigned long long key = sk[q].key(i);
      const bool ok = (uint32_t)i < cnt[q];
      const uint32_t pos = atomicAdd(s_cnt + (ok ? (uint32_t)(key >> dsh) : dummy), 1u);
      if (ok) *reinterpret_cast<uint2 *>(stage + (size_t)pos * 2) = make_uint2((uint32_t)(key >> 32), (uint32_t)key);
    }
  }
  __syncthreads();
  const uint2 *st2 = reinterpret_cast<const uint2 *>(stage);
  uint2 *out2 = reinterpret_cast<uint2 *>(out);
  const int dsh32 = 32 - nbits;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const uint32_t j = (uint32_t)(q * NT + tid);
    if (j < total) {
      const uint2 v = st2[j];
      const long long gd = s_gd[v.x >> dsh32];
      if (gd != kDropRun) out2[gd + (long long)j] = v;
    }
  }
}

// ------------------------------------------------------------------ host side
static int skm_ceil_log2(double x) {
  int b = 0;
  while ((double)(1ull << b) < x && b < 62) ++b;
  return b;
}

void dev_skm_scatter(Ctx &c, const ReadsView &r, int k, int n_dst, const uint64_t *dst_ptrs, const int64_t *dst_caps, int64_t stride,
                     int64_t *counts_out) {
  if (!skm_supported(k)) throw std::invalid_argument("super-k-mer records need 16 <= k <= 26");
  if (n_dst < 1 || n_dst > kSkmMaxDst) throw std::invalid_argument("skm_scatter: 1..16 destinations");
  if (stride < 1) stride = 1;
  const bool count_only = dst_ptrs == nullptr;
  const int K1 = k + 1;
  const uint32_t *sbits = dev_start_bits(c, r);
  c.slab_reserve(1 << 20);
  SkmArgs a;
  a.n_dst = n_dst;
  a.cmax = skm_cmax(K1);
  const int nm = K1 >= 21 ? 14 : 8;   // m-mers per window; m = 8..14 (K1 >= 21) or 10..13
  a.m = K1 - nm + 1;
  for (int d = 0; d < kSkmMaxDst; ++d) {
    a.dst[d] = (!count_only && d < n_dst) ? dst_ptrs[d] : 0ull;
    a.cap[d] = (!count_only && d < n_dst) ? (unsigned long long)std::max<int64_t>(dst_caps[d], 0) : 0ull;
  }
  a.cursor = c.alloc<unsigned long long>(2 * kSkmMaxDst);
  MF_CUDA(cudaMemsetAsync(a.cursor, 0, sizeof(unsigned long long) * 2 * kSkmMaxDst, c.stream));
  const int64_t tiles_all = div_ceil64(r.n_bases, (int64_t)kSkmNT * 16), ntiles = div_ceil64(tiles_all, stride);
  if (ntiles > 0) {
    Stage st(c, count_only ? "skm_sample" : "skm_scatter");
    const size_t smem = skm_scatter_smem_bytes(k, count_only);
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)c.sm_count * 2 * 8);
    ReadsSrc src{r.packed, sbits, r.n_bases, k};
#define MF_SKM_LAUNCH(NMv, COv)                                                 \
  {                                                                             \
    auto kern = k_skm_scatter<NMv, COv>;                                        \
    skm_set_smem(kern, smem);                                                   \
    kern<<<grid, kSkmNT, smem, c.stream>>>(src, a, ntiles, stride);             \
  }
    if (nm == 14) {
      if (count_only) MF_SKM_LAUNCH(14, true) else MF_SKM_LAUNCH(14, false)
    } else {
      if (count_only) MF_SKM_LAUNCH(8, true) else MF_SKM_LAUNCH(8, false)
    }
#undef MF_SKM_LAUNCH
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  std::vector<unsigned long long> cur(2 * kSkmMaxDst);
  c.d2h(cur.data(), a.cursor, sizeof(unsigned long long) * 2 * kSkmMaxDst);
  for (int d = 0; d < n_dst; ++d) {
    counts_out[d] = (int64_t)cur[d];
    counts_out[n_dst + d] = (int64_t)cur[n_dst + d];
  }
}

void dev_count_skm(Ctx &c, const uint64_t *recs, const int64_t *chunk_start, const int64_t *chunk_size, int n_chunks, int64_t n_keys,
                   int k, int min_count, uint32_t *keys, uint32_t *scratch, int64_t capacity, EdgesView *out) {
  if (!skm_supported(k)) throw std::invalid_argument("super-k-mer records need 16 <= k <= 26");
  if (min_count < 1) throw std::invalid_argument("min_count must be >= 1");
  const int K1 = k + 1, key_bits = 2 * K1;
  int64_t n_rec = 0;
  for (int i = 0; i < n_chunks; ++i) {
    if (chunk_size[i] < 0 || chunk_start[i] < 0) throw std::invalid_argument("count_skm: negative chunk");
    n_rec += chunk_size[i];
  }
  if (n_rec == 0 || n_keys == 0) {
    const int64_t z = 0;
    const int32_t zs = 0;
    dev_count_finish(c, keys, scratch, 0, &z, &z, &zs, 0, 1, k, 1, min_count, out, nullptr);
    return;
  }
  if (n_keys < n_rec || n_keys > n_rec * skm_slots(K1)) throw std::invalid_argument("count_skm: n_keys does not fit the record count");
  // level-1 width as count_plan chooses it for the streamed finish (engine.cu): bins of ~12 M keys, 512 at 5 Gbp
  const int l1_bits = std::max(1, std::min({10, key_bits, skm_ceil_log2((double)n_keys / 1.2e7)}));
  const int nb1 = 1 << l1_bits;
  // tiles of the chunks: 16 key slots per thread, i.e. 2 records of <= 8 keys or 4 records of <= 4
  const int slots = skm_slots(K1), rec_tile = kSkmTileKeys / slots;
  std::vector<int64_t> cs, cz, tb;
  std::vector<int32_t> cg;
  tb.push_back(0);
  for (int i = 0; i < n_chunks; ++i) {
    if (chunk_size[i] == 0) continue;
    cs.push_back(chunk_start[i]);
    cz.push_back(chunk_size[i]);
    cg.push_back(0);
    tb.push_back(tb.back() + div_ceil64(chunk_size[i], rec_tile));
  }
  const int nch = (int)cs.size();
  const int64_t ntiles = tb.back();
  c.slab_reserve((size_t)ntiles * sizeof(TileDesc) + (size_t)nch * 64 + (size_t)nb1 * 64 + (1 << 20));
  int64_t *d_cs = c.alloc<int64_t>(nch), *d_cz = c.alloc<int64_t>(nch), *d_tb = c.alloc<int64_t>(nch + 1);
  int32_t *d_cg = c.alloc<int32_t>(nch);
  TileDesc *d_tiles = c.alloc<TileDesc>(ntiles);
  unsigned long long *d_hist = c.alloc<unsigned long long>(nb1), *d_cursor = c.alloc<unsigned long long>(nb1),
                     *d_limit = c.alloc<unsigned long long>(nb1);
  MF_CUDA(cudaMemcpyAsync(d_cs, cs.data(), sizeof(int64_t) * nch, cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemcpyAsync(d_cz, cz.data(), sizeof(int64_t) * nch, cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemcpyAsync(d_cg, cg.data(), sizeof(int32_t) * nch, cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemcpyAsync(d_tb, tb.data(), sizeof(int64_t) * (nch + 1), cudaMemcpyHostToDevice, c.stream));
  k_build_tiles<<<(unsigned)div_ceil64(ntiles, 256), 256, 0, c.stream>>>(ChunkTable{d_cs, d_cz, d_cg, d_tb, nch}, rec_tile, ntiles, d_tiles);
  MF_LAUNCH_CHECK();
  c.launches++;
  const unsigned long long *urecs = reinterpret_cast<const unsigned long long *>(recs);
  std::vector<unsigned long long> start(nb1), limit(nb1), cur(nb1), hs(nb1);
  const char *env_s = getenv("MFSDBG_SAMPLED_STRIDE");
  const int64_t stride0 = (ntiles >= 4096 && !(getenv("MFSDBG_SAMPLED_HIST") && atoi(getenv("MFSDBG_SAMPLED_HIST")) == 0))
                              ? std::max(1, env_s && *env_s ? atoi(env_s) : 64) : 1;
  bool done = false;
  for (int attempt = 0; attempt < 2 && !done; ++attempt) {
    const int64_t stride = attempt == 0 ? stride0 : 1;
    if (attempt == 1 && stride0 == 1) break;
    MF_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nb1, c.stream));
    {
      Stage st(c, "skm_hist");
      const size_t smem = ((size_t)nb1 + 32) * 4;
      auto hk = slots == 4 ? k_skm_hist<4> : k_skm_hist<8>;
      skm_set_smem(hk, smem);
      const int64_t nsamp = div_ceil64(ntiles, stride);
      hk<<<(unsigned)std::min<int64_t>(nsamp, (int64_t)c.sm_count * 4), kSkmNT, smem, c.stream>>>(urecs, d_tiles, ntiles, stride, K1,
                                                                                                        l1_bits, d_hist);
      MF_LAUNCH_CHECK();
      c.launches++;
    }
    c.d2h(hs.data(), d_hist, sizeof(unsigned long long) * nb1);
    // a strided sample of tiles: scale by the records it saw (the last tile of a chunk is short)
    unsigned long long acc = 0;
    if (stride == 1) {
      for (int b = 0; b < nb1; ++b) {
        start[b] = acc;
        acc += hs[b];
        acc = (acc + 1ull) & ~1ull;
        limit[b] = acc;
      }
    } else {
      unsigned long long seen = 0;
      for (int b = 0; b < nb1; ++b) seen += hs[b];
      const double scale = seen ? (double)n_keys / (double)seen : 1.0;
      for (int b = 0; b < nb1; ++b) {
        start[b] = acc;
        acc += (unsigned long long)((double)hs[b] * scale * 1.10) + 65536ull;
        acc = (acc + 1ull) & ~1ull;
        limit[b] = acc;
      }
    }
    if ((int64_t)acc > capacity) {
      if (stride == 1) throw std::invalid_argument("count_skm: key buffers hold " + std::to_string(capacity) + " records, " + std::to_string(acc) + " needed");
      continue;   // the slack does not fit: exact regions
    }
    c.h2d(d_cursor, start.data(), sizeof(unsigned long long) * nb1);
    c.h2d(d_limit, limit.data(), sizeof(unsigned long long) * nb1);
    {
      Stage st(c, "skm_l1_scatter");
      const size_t smem = skm_kscatter_smem_bytes(l1_bits);
      const int bpt = std::max(1, nb1 / kSkmNT);
      auto kern = slots == 4 ? (bpt == 1 ? k_skm_kscatter<1, 4> : k_skm_kscatter<2, 4>) : (bpt == 1 ? k_skm_kscatter<1, 8> : k_skm_kscatter<2, 8>);
      skm_set_smem(kern, smem);
      kern<<<(unsigned)ntiles, kSkmNT, smem, c.stream>>>(urecs, d_tiles, K1, l1_bits, d_cursor, d_limit, keys);
      MF_LAUNCH_CHECK();
      c.launches++;
    }
    c.d2h(cur.data(), d_cursor, sizeof(unsigned long long) * nb1);
    done = true;
    for (int b = 0; b < nb1; ++b) done = done && cur[b] <= limit[b];
    if (!done) {
      if (stride == 1) throw std::runtime_error("count_skm: the exact histogram and the scatter disagree");
      Stage mark(c, "sampled_overflow");
    }
  }
  if (!done) throw std::runtime_error("count_skm: level 1 did not complete");
  std::vector<int64_t> fs, fz;
  std::vector<int32_t> fg;
  int64_t total = 0;
  for (int b = 0; b < nb1; ++b) {
    fs.push_back((int64_t)start[b]);
    fz.push_back((int64_t)(cur[b] - start[b]));
    fg.push_back(b);
    total += fz.back();
  }
  if (total != n_keys) throw std::runtime_error("count_skm: the records held " + std::to_string(total) + " keys, the senders announced " + std::to_string(n_keys));
  dev_count_finish(c, keys, scratch, total, fs.data(), fz.data(), fg.data(), nb1, nb1, k, l1_bits, min_count, out, nullptr);
  if (out->n_edges > 0) {   // the count ran on mixed keys: map the solid edges back (their order stays the mixed one)
    Stage st(c, "skm_unmix");
    k_skm_unmix_edges<<<(unsigned)div_ceil64(out->n_edges, 256), 256, 0, c.stream>>>(const_cast<uint32_t *>(out->edges), out->n_edges, out->words, K1);
    MF_LAUNCH_CHECK();
    c.launches++;
    MF_CUDA(cudaStreamSynchronize(c.stream));
  }
}

}  // namespace mf
