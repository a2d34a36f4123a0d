// count_stream.cuh -- pieces shared by the streamed count finishes (k_count_stream2 for keys of 33..64 bits, k_count_stream_w
// for wider ones): the mbarrier / bulk-copy primitives (cp.async.bulk + mbarrier: SASS UBLKCP + SYNCS), the split of the
// bucket table into per-CTA ranges, the multi-pass kernel for buckets whose distinct keys crowd the shared table, the
// distinct-ratio probe that sizes the buckets, and the TMA-fed scatter of a partition level.
#pragma once
#include "common.cuh"
#include "local.cuh"
#include "partition.cuh"

namespace mf {

constexpr int kCsNT = 512;
constexpr int kCsSlotsLog = 12;
constexpr int kCsSlots = 1 << kCsSlotsLog;   // table slots
constexpr int kCsSolidMax = 768;             // distinct solid keys of one bucket on this path
constexpr int kCsWin = 256;                  // bucket boundaries held in shared memory at a time
constexpr int kCsProbeLimit = 64;
constexpr int kCsArenaBlock = 4096;          // edges reserved per global atomic (>= kCsSolidMax)
constexpr int kCsPairsMax = 128;             // solid keys ranked by all-pairs comparison; more take a counting split

// ---- mbarrier / bulk-copy primitives (sm_90+ PTX; SASS: UBLKCP + SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  // a warp that finds the stage not ready backs off for a few dozen ns instead of spinning at full issue rate: under ncu the
  // spin loop alone was 8 % of all instructions the count kernel executed
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(40);
  }
}

// the producer's wait: it has nothing else to do, so it polls rarely (its spin was 9 % of the count kernel's instructions)
__device__ __forceinline__ void mbar_wait_slow(unsigned long long *bar, uint32_t parity) {
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(400);
  }
}

inline size_t count_multipass_smem_bytes() {
  // tkeys u64[4096] | skeys u64[768] | tcnt u32[4096] | scnt u32[768] | bins u32[260] | small u32[64] | scratch u32[40]
  // | flag i32[16] | permA,permB,rk u16[768]
  return (size_t)kCsSlots * 8 + (size_t)kCsSolidMax * 8 + (size_t)kCsSlots * 4 + (size_t)kCsSolidMax * 4 + 260 * 4 + 64 * 4 + 40 * 4 +
         16 * 4 + 3 * (size_t)kCsSolidMax * 2;
}

// cta_first[g] = first bucket slot of CTA g's range: ranges hold equal shares of the keys (bkt_start is monotone)
__global__ void k_split_ranges(const int64_t *bkt_start, const int64_t *bkt_size, int nslots, int G, int32_t *cta_first) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > G) return;
  if (g == G) { cta_first[G] = nslots; return; }
  const int64_t s0 = bkt_start[0], total = bkt_start[nslots - 1] + bkt_size[nslots - 1] - s0;
  const int64_t target = s0 + (int64_t)(((__int128)total * g) / G);
  int lo = 0, hi = nslots;   // first slot with bkt_start >= target
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bkt_start[mid] < target) lo = mid + 1; else hi = mid;
  }
  cta_first[g] = g == 0 ? 0 : lo;
}

// One CTA per bucket whose DISTINCT keys crowd the streamed kernel's table (k_count_stream2 puts it on the bail list): the
// bucket's key range is cut into 2^passes_log equal sub-ranges and the bucket is re-read once per sub-range and sweep (sweep 1
// totals the solid keys so that the arena space is reserved in one piece, sweep 2 emits); keys outside the pass's sub-range are
// skipped -- the table only ever holds a fraction of the distinct keys, the output stays sorted.
__global__ void __launch_bounds__(kCsNT, 2) k_count_multipass(LocalArgs a, int passes_log) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr int NT = kCsNT;
  unsigned long long *tkeys = reinterpret_cast<unsigned long long *>(smraw);
  unsigned long long *skeys = tkeys + kCsSlots;
  uint32_t *tcnt = reinterpret_cast<uint32_t *>(skeys + kCsSolidMax);
  uint32_t *scnt = tcnt + kCsSlots;
  uint32_t *bins = scnt + kCsSolidMax;
  uint32_t *s_small = bins + 260;
  uint32_t *scratch = s_small + 64;
  int *s_flag = reinterpret_cast<int *>(scratch + 40);   // 0 crowded, 1 ok, 2..3 arena base, 4 ns, 8 bail, 9 total, 10 overflow
  uint16_t *permA = reinterpret_cast<uint16_t *>(s_flag + 16);
  uint16_t *permB = permA + kCsSolidMax, *rk = permB + kCsSolidMax;
  uint32_t *whist32 = tcnt;   // sub-bin counters [1025] of the many-solid-keys sort alias the (swept, empty) counts

  const int tid = threadIdx.x;
  const uint32_t m = (uint32_t)a.min_count;
  const int We = a.words_edge;
  for (int i = tid; i < kCsSlots; i += NT) {
    tkeys[i] = kEmptyKey;
    tcnt[i] = 0u;
  }
  if (tid < 64) s_small[tid] = 0;
  if (tid < 16) s_flag[tid] = 0;
  __syncthreads();

  auto insert = [&](unsigned long long key) {
    uint32_t x = (uint32_t)key ^ ((uint32_t)(key >> 32) * 0x9E3779B1u);
    uint32_t h = (x * 0x85EBCA6Bu) >> (32 - kCsSlotsLog);
    for (int probe = 0; probe < kCsProbeLimit; ++probe) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tkeys + h);
      if (cur == kEmptyKey) cur = atomicCAS(tkeys + h, kEmptyKey, key);
      if (cur == kEmptyKey || cur == key) {
        atomicAdd(tcnt + h, 1u);
        return;
      }
      h = (h + 1) & (kCsSlots - 1);
    }
    s_flag[0] = 1;   // a sub-range still crowds the table: the general path takes the bucket
  };
  // ---- pass end: sweep + clear, solid keys -> ordered edge records.  Called by all threads, after a __syncthreads.
  // mode 1: count sweep (only totals the solid keys in s_flag[9], failure -> s_flag[8]); mode 2: emit sweep (writes at the base
  // held in s_flag[2..3] and advances it)
  auto finish_pass = [&](int mode) {
    const bool crowded = s_flag[0] != 0;
    for (int h = tid; h < kCsSlots; h += NT) {
      const uint32_t c = tcnt[h];
      if (c) {
        const unsigned long long key = tkeys[h];
        tkeys[h] = kEmptyKey;
        tcnt[h] = 0u;
        if (a.counting && mode != 1 && c < 64u) atomicAdd(s_small + c, 1u);
        if (c >= m) {
          const int q = atomicAdd(s_flag + 4, 1);
          if (q < kCsSolidMax) {
            skeys[q] = key;
            scnt[q] = c;
          }
        }
      }
    }
    __syncthreads();
    const int ns_raw = s_flag[4];
    const bool bail = crowded || ns_raw > kCsSolidMax;
    const uint32_t ns = bail ? 0u : (uint32_t)ns_raw;
    if (a.counting && mode != 1) {
      // multiplicity histogram of the distinct keys (<prefix>.counting): small counts from shared memory, counts >= 64 are
      // solid keys (the host routes --min-count > 64 with a histogram request to the general kernel)
      if (tid < 64) {
        const uint32_t v = s_small[tid];
        s_small[tid] = 0;
        if (v && !bail) atomicAdd(a.counting + tid, (unsigned long long)v);
      }
      for (uint32_t q = tid; q < ns; q += NT) {
        const uint32_t c = scnt[q];
        if (c >= 64u) atomicAdd(a.counting + (c > (uint32_t)kMaxMul ? (uint32_t)kMaxMul : c), 1ull);
      }
    }
    if (tid == 0) {
      if (mode == 1) {
        if (bail) s_flag[8] = 1;
        s_flag[9] += (int)ns;
        s_flag[1] = 0;
      } else {
        s_flag[1] = ns > 0 && !s_flag[10];   // the base in s_flag[2..3] is advanced after the write below
      }
    }
    if (ns > 1 && ns <= (uint32_t)kCsPairsMax)
      for (uint32_t q = tid; q < ns; q += NT) bins[q] = 0;
    __syncthreads();
    if (s_flag[1]) {
      const uint16_t *cur = nullptr;
      if (ns > 1 && ns <= (uint32_t)kCsPairsMax) {
        // keys are distinct: rank = number of smaller keys; the comparisons of one key are split over NT / nsp threads
        uint32_t nsp = 32;
        while (nsp < ns) nsp <<= 1;
        const uint32_t parts = NT / nsp, q = tid & (nsp - 1), part = tid / nsp;
        if (q < ns) {
          const uint32_t per = (ns + parts - 1) / parts;
          const uint32_t o0 = part * per, o1 = min(ns, o0 + per);
          const unsigned long long kq = skeys[q];
          uint32_t r = 0;
          for (uint32_t o = o0; o < o1; ++o) r += skeys[o] < kq;
          if (r) atomicAdd(bins + q, r);
        }
        __syncthreads();
        for (uint32_t q2 = tid; q2 < ns; q2 += NT) permA[bins[q2]] = (uint16_t)q2;
        __syncthreads();
        cur = permA;
      } else if (ns > (uint32_t)kCsPairsMax) {
        // many solid keys: one counting split on the 10 bits below the keys' common range, then a rank fix inside each sub-bin
        unsigned long long *s_mm = reinterpret_cast<unsigned long long *>(scratch);   // [0] min, [1] max (8-byte aligned)
        if (tid == 0) { s_mm[0] = ~0ull; s_mm[1] = 0ull; }
        for (int i = tid; i <= 1024; i += NT) whist32[i] = 0u;
        __syncthreads();
        {
          unsigned long long mn = ~0ull, mx = 0ull;
          for (uint32_t q = tid; q < ns; q += NT) { const unsigned long long kq = skeys[q]; mn = min(mn, kq); mx = max(mx, kq); }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          }
          if ((tid & 31) == 0) { atomicMin(s_mm, mn); atomicMax(s_mm + 1, mx); }
        }
        __syncthreads();
        const unsigned long long kmin = s_mm[0];
        const int span_bits = 64 - __clzll((long long)((s_mm[1] - kmin) | 1ull));
        const int sh = span_bits > 10 ? span_bits - 10 : 0;
        for (uint32_t q = tid; q < ns; q += NT) rk[q] = (uint16_t)atomicAdd(whist32 + (uint32_t)((skeys[q] - kmin) >> sh), 1u);
        __syncthreads();
        block_excl_scan<NT>(whist32, 1025, scratch + 4);
        for (uint32_t q = tid; q < ns; q += NT) permB[whist32[(uint32_t)((skeys[q] - kmin) >> sh)] + rk[q]] = (uint16_t)q;
        __syncthreads();
        for (uint32_t p2 = tid; p2 < ns; p2 += NT) {
          const uint32_t q = permB[p2];
          const unsigned long long kq = skeys[q];
          const uint32_t d = (uint32_t)((kq - kmin) >> sh);
          const uint32_t b2 = whist32[d], e2 = whist32[d + 1];
          uint32_t r = 0;
          for (uint32_t o = b2; o < e2; ++o) r += skeys[permB[o]] < kq;
          permA[b2 + r] = (uint16_t)q;
        }
        __syncthreads();
        for (int i = tid; i <= 1024; i += NT) whist32[i] = 0u;   // give the counts back clean
        cur = permA;
      }
      const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
      for (uint32_t q = tid; q < ns; q += NT) {
        const uint32_t e = cur ? cur[q] : q;
        const unsigned long long key = skeys[e];
        const uint32_t kw[2] = {(uint32_t)(key >> 32), (uint32_t)key};
        write_edge<2>(a.arena + (base + q) * (unsigned long long)We, kw, We, scnt[e]);
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (mode == 2) {
        const unsigned long long nb2 = (((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2]) + ns;
        s_flag[2] = (int)(uint32_t)nb2;
        s_flag[3] = (int)(uint32_t)(nb2 >> 32);
      }
      s_flag[0] = 0;
      s_flag[4] = 0;
    }
    // the barrier the caller issues next publishes the reset
  };

  const WorkItem wi = a.work[blockIdx.x];
  const int slot = wi.slot;
  const int64_t n = a.bkt_size[slot];
  const uint2 *keys = reinterpret_cast<const uint2 *>(a.in) + a.bkt_start[slot];
  // key range of the bucket
  unsigned long long *s_mm = reinterpret_cast<unsigned long long *>(scratch);
  if (tid == 0) { s_mm[0] = ~0ull; s_mm[1] = 0ull; }
  __syncthreads();
  {
    unsigned long long mn = ~0ull, mx = 0ull;
    for (int64_t i = tid; i < n; i += NT) {
      const uint2 v = keys[i];
      const unsigned long long key = ((unsigned long long)v.x << 32) | v.y;
      mn = min(mn, key);
      mx = max(mx, key);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) { atomicMin(s_mm, mn); atomicMax(s_mm + 1, mx); }
  }
  __syncthreads();
  const unsigned long long kmin = s_mm[0];
  const int span_bits = 64 - __clzll((long long)((s_mm[1] - kmin) | 1ull));
  const int sh = span_bits > passes_log ? span_bits - passes_log : 0;
  const uint32_t npass = (uint32_t)((s_mm[1] - kmin) >> sh) + 1u;
  __syncthreads();
  for (int sweep = 1; sweep <= 2; ++sweep) {
    for (uint32_t ps = 0; ps < npass; ++ps) {
      for (int64_t i = tid; i < n; i += NT) {
        const uint2 v = keys[i];
        const unsigned long long key = ((unsigned long long)v.x << 32) | v.y;
        if ((uint32_t)((key - kmin) >> sh) == ps) insert(key);
      }
      __syncthreads();
      finish_pass(sweep);
      __syncthreads();
    }
    if (sweep == 1) {
      if (tid == 0) {
        if (s_flag[8]) {   // a sub-range still crowds the table: the general path takes the bucket
          const int p = atomicAdd(a.bail_count, 1);
          a.bail_list[p] = slot;
        } else {
          const unsigned long long tot = (unsigned long long)(uint32_t)s_flag[9];
          const unsigned long long pos = atomicAdd(a.arena_cursor, tot);
          const int ok = pos + tot <= a.arena_cap;
          if (!ok) atomicExch(a.overflow_flag, 1);
          a.desc_off[slot] = (int64_t)pos;
          a.desc_cnt[slot] = ok ? (int64_t)tot : 0;
          s_flag[2] = (int)(uint32_t)pos;
          s_flag[3] = (int)(uint32_t)(pos >> 32);
          // arena overflow: the host redoes the finish with a larger arena and WITHOUT the histogram (a retry must not
          // count twice), so sweep 2 still runs -- for the histogram only
          if (!ok) s_flag[10] = 1;
        }
      }
      __syncthreads();
      if (s_flag[8]) return;
      if (s_flag[10] && !a.counting) return;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// distinct-ratio probe: for a few whole prefix ranges ("samples", each a list of chunks of one segment) the keys whose
// next `fbits` bits equal `pattern` form what a final bucket of that range would hold; count them and their distinct keys
// in a small global table per sample.
struct ProbeChunk {
  int64_t start, size;
  int32_t sample;
  int32_t fbits;   // filter width
};
constexpr int kProbeSlots = 1 << 14;
__global__ void k_probe_distinct(const uint32_t *__restrict__ keys, const ProbeChunk *__restrict__ chunks, int bit_off, uint32_t pattern,
                                 unsigned long long *tables /*[samples][kProbeSlots]*/, unsigned long long *stats /*[samples][2]*/) {
  const ProbeChunk ch = chunks[blockIdx.y];
  const uint2 *src = reinterpret_cast<const uint2 *>(keys) + ch.start;
  unsigned long long *tab = tables + (size_t)ch.sample * kProbeSlots;
  const uint32_t want = ch.fbits ? (pattern & ((1u << ch.fbits) - 1u)) : 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ch.size; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 v = src[i];
    const uint32_t r[2] = {v.x, v.y};
    if (ch.fbits && rec_digit<2>(r, bit_off, ch.fbits) != want) continue;
    const unsigned long long key = ((unsigned long long)v.x << 32) | v.y;
    atomicAdd(stats + 2 * ch.sample, 1ull);
    uint32_t h = (uint32_t)((key * 0x9e3779b97f4a7c15ull) >> (64 - 14));
    for (int probe = 0; probe < kProbeSlots; ++probe) {
      unsigned long long cur = tab[h];
      if (cur == kEmptyKey) cur = atomicCAS(tab + h, kEmptyKey, key);
      if (cur == kEmptyKey) { atomicAdd(stats + 2 * ch.sample + 1, 1ull); break; }
      if (cur == key) break;
      h = (h + 1) & (kProbeSlots - 1);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// TMA-fed scatter of one partition level over 2-word records: the tile arrives in shared memory with one bulk copy (no
// per-thread loads, no records held in registers), is counted and re-read there, staged in bin order and copied out
// coalesced.  KPT keys per thread: 12 (6144-key tiles) with up to 1024 bins, 10 with 2048 -- two CTAs per SM either way.
template <int KPT, int NT = 512>
inline size_t scatter_tma_smem_bytes(int nbits) {
  const size_t nb = (size_t)1 << nbits, T = (size_t)NT * KPT;
  return (T + 4) * 8 + T * 8 + (nb + 32) * 4 + nb * 8 + 48 * 4 + 16;
}
// lean digit of a 2-word record for bit_off < 32: dg = (x * nb) >> xbits with x = the xbits bits at bit_off; the bit-prefix
// levels pass nb = 2^xbits, so one formula serves both (these kernels are instruction-issue bound: 37 -> ~12 per record)
struct Digit2 {
  int sh, shx, xbits;
  uint32_t nb;
  __device__ __forceinline__ Digit2(const LevelArgs &a, uint32_t seg_nb) {
    xbits = seg_nb ? a.xbits : a.nbits;
    nb = seg_nb ? seg_nb : (1u << a.nbits);
    sh = a.bit_off;
    shx = 32 - xbits;
  }
  __device__ __forceinline__ uint32_t operator()(uint2 v) const { return ((__funnelshift_l(v.y, v.x, sh) >> shx) * nb) >> xbits; }
};

template <int NT, int KPT, int BPT>
__global__ void __launch_bounds__(NT, KPT <= 7 ? 3 : 2) k_scatter_tma(const uint32_t *__restrict__ in, const TileDesc *__restrict__ tiles, int64_t ntiles,
                                                      LevelArgs a, unsigned long long *__restrict__ cursor, uint32_t *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr int T = NT * KPT;
  const int nbins = 1 << a.nbits;
  const int tid = threadIdx.x;
  uint2 *X = reinterpret_cast<uint2 *>(smraw);                       // [T + 4]
  uint2 *Y = X + T + 4;                                              // [T]
  long long *s_gd = reinterpret_cast<long long *>(Y + T);            // [nbins]
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_gd + nbins);      // [nbins + 32]
  uint32_t *scratch = s_cnt + nbins + 32;                            // [48]
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(scratch + 48);

  // persistent: CTA b takes tiles b, b + grid, ...; the bulk copy of the next tile is issued as soon as the staging pass
  // has emptied X, so it flies behind the copy-out of the current tile
  int64_t t = blockIdx.x;
  if (t >= ntiles) return;
  TileDesc d = tiles[t];
  auto issue = [&](const TileDesc &td) {   // thread 0 only
    const int o = (int)(td.base & 1);
    const uint32_t bytes = (uint32_t)(((td.n + o) * 8 + 15) & ~15);
    mbar_expect_tx(mbar, bytes);
    bulk_g2s(X, reinterpret_cast<const uint2 *>(in) + (td.base - o), bytes, mbar);
  };
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
  __syncthreads();
  if (tid == 0) issue(d);
  for (uint32_t it = 0;; ++it) {
    const int64_t tn = t + gridDim.x;
    const bool more = tn < ntiles;
    TileDesc dn = d;
    if (more) dn = tiles[tn];   // on its way while this tile is processed
    const int n = d.n;
    const Digit2 digit(a, a.seg_nb ? a.seg_nb[d.seg] : 0u);
    const uint2 *Xo = X + (int)(d.base & 1);
    mbar_wait(mbar, it & 1u);
    // phase A: one shared atomic per record; its return value is the record's rank inside its bin, kept with the digit in a
    // register so that the staging pass needs neither a second atomic nor the digit again
    uint32_t rk[KPT];
    if (n == T) {
#pragma unroll
      for (int q = 0; q < KPT; ++q) {
        const uint32_t dg = digit(Xo[q * NT + tid]);
        rk[q] = (dg << 16) | atomicAdd(s_cnt + dg, 1u);
      }
    } else {
#pragma unroll
      for (int q = 0; q < KPT; ++q) {
        const int j = q * NT + tid;
        rk[q] = 0xffffffffu;
        if (j < n) {
          const uint32_t dg = digit(Xo[j]);
          rk[q] = (dg << 16) | atomicAdd(s_cnt + dg, 1u);
        }
      }
    }
    __syncthreads();
    const uint32_t total = bins_scan_reserve<NT, BPT>(s_cnt, s_gd, scratch, cursor + (size_t)d.seg * nbins, nbins);
    // phase B: stage every record at its bin's start + rank
#pragma unroll
    for (int q = 0; q < KPT; ++q)
      if (rk[q] != 0xffffffffu) Y[s_cnt[rk[q] >> 16] + (rk[q] & 0xffffu)] = Xo[q * NT + tid];
    __syncthreads();   // X is free, Y is complete, the bin starts in s_cnt are dead
    if (tid == 0 && more) issue(dn);
    for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
    uint2 *out2 = reinterpret_cast<uint2 *>(out);
#pragma unroll
    for (int q = 0; q < KPT; ++q) {   // fully unrolled: all the shared loads of a thread are in flight together
      const uint32_t j = (uint32_t)(q * NT + tid);
      if (j < total) {
        const uint2 v = Y[j];
        out2[s_gd[digit(v)] + (long long)j] = v;
      }
    }
    if (!more) break;
    __syncthreads();   // Y and s_gd are free, the cleared counters are visible
    t = tn;
    d = dn;
  }
}


// The same kernel for records of W = 3..10 words (keys of k >= 32, sdbg items of k >= 23): the tile lives in shared memory,
// so its size no longer depends on how many records a thread can hold in registers (the register-staged k_level_scatter
// gets 1024-2048 records per tile at these widths -- one or two records per bin and run).  T = NT * KPT records, X and Y of
// 48 KB each; the digit comes from the first two words (bit_off < 32).
template <int W>
struct ScatterWCfg {
  static constexpr int NT = 512;
  static constexpr int KPT = (48 * 1024 / (4 * W)) / NT < 1 ? 1 : (48 * 1024 / (4 * W)) / NT;   // 8, 6, 4, 4, 3, 3, 2, 2
  static constexpr int T = NT * KPT;
  static constexpr int AL = (W % 4 == 0) ? 1 : ((W % 2 == 0) ? 2 : 4);   // records per 16-byte boundary
};
template <int W>
inline size_t scatter_tma_w_smem_bytes(int nbits) {
  using C = ScatterWCfg<W>;
  const size_t nb = (size_t)1 << nbits;
  return ((size_t)(C::T + 4) * W * 4 + 15) / 16 * 16 + (size_t)C::T * W * 4 + nb * 8 + (nb + 32) * 4 + 48 * 4 + 16 + 16;
}
template <int W, int BPT>
__global__ void __launch_bounds__(512, 2) k_scatter_tma_w(const uint32_t *__restrict__ in, const TileDesc *__restrict__ tiles, int64_t ntiles,
                                                         LevelArgs a, unsigned long long *__restrict__ cursor, uint32_t *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smraw[];
  using C = ScatterWCfg<W>;
  constexpr int NT = C::NT, KPT = C::KPT, T = C::T, AL = C::AL;
  const int nbins = 1 << a.nbits;
  const int tid = threadIdx.x;
  uint32_t *X = reinterpret_cast<uint32_t *>(smraw);                                                   // [(T + 4) * W]
  uint32_t *Y = X + (((size_t)(T + 4) * W + 3) & ~(size_t)3);                                          // [T * W]
  long long *s_gd = reinterpret_cast<long long *>(Y + (size_t)T * W + (((size_t)T * W) & 1));          // [nbins]
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_gd + nbins);                                        // [nbins + 32]
  uint32_t *scratch = s_cnt + nbins + 32;                                                              // [48]
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(scratch + 48 + ((nbins + 32 + 48) & 1));

  int64_t t = blockIdx.x;
  if (t >= ntiles) return;
  TileDesc d = tiles[t];
  auto issue = [&](const TileDesc &td) {   // thread 0 only
    const int o = (int)(td.base % AL);
    const uint32_t bytes = (uint32_t)((((td.n + o) * W * 4) + 15) & ~15);
    mbar_expect_tx(mbar, bytes);
    bulk_g2s(X, in + (td.base - o) * W, bytes, mbar);
  };
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
  __syncthreads();
  if (tid == 0) issue(d);
  for (uint32_t it = 0;; ++it) {
    const int64_t tn = t + gridDim.x;
    const bool more = tn < ntiles;
    TileDesc dn = d;
    if (more) dn = tiles[tn];
    const int n = d.n;
    const Digit2 digit(a, a.seg_nb ? a.seg_nb[d.seg] : 0u);
    const uint32_t *Xo = X + (size_t)(d.base % AL) * W;
    mbar_wait(mbar, it & 1u);
    uint32_t rk[KPT];
#pragma unroll
    for (int q = 0; q < KPT; ++q) {
      const int j = q * NT + tid;
      rk[q] = 0xffffffffu;
      if (j < n) {
        const uint32_t dg = digit(make_uint2(Xo[(size_t)j * W], Xo[(size_t)j * W + 1]));
        rk[q] = (dg << 16) | atomicAdd(s_cnt + dg, 1u);
      }
    }
    __syncthreads();
    const uint32_t total = bins_scan_reserve<NT, BPT>(s_cnt, s_gd, scratch, cursor + (size_t)d.seg * nbins, nbins);
#pragma unroll
    for (int q = 0; q < KPT; ++q) {
      if (rk[q] != 0xffffffffu) {
        const uint32_t *src = Xo + (size_t)(q * NT + tid) * W;
        uint32_t *dst = Y + (size_t)(s_cnt[rk[q] >> 16] + (rk[q] & 0xffffu)) * W;
        uint32_t rec[W];
#pragma unroll
        for (int c = 0; c < W; ++c) rec[c] = src[c];
        stage_store<W>(dst, rec);
      }
    }
    __syncthreads();   // X is free, Y is complete
    if (tid == 0 && more) issue(dn);
    for (int i = tid; i < nbins + 32; i += NT) s_cnt[i] = 0;
    // W / VEC threads per record, consecutive threads write consecutive words (a thread per record wrote W words 4 W bytes apart)
    staged_copy_out<W, NT>(Y, s_gd, total, a, out, [&](const uint32_t *r) { return digit(make_uint2(r[0], r[1])); });
    if (!more) break;
    __syncthreads();
    t = tn;
    d = dn;
  }
}

}  // namespace mf
