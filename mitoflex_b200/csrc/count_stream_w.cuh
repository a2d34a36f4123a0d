// count_stream_w.cuh -- the streamed count finish for keys of more than 64 bits (k >= 32, W = 3..10 words).
//
// Same shape as k_count_stream (count_stream.cuh): persistent CTAs stream contiguous runs of buckets through a
// cp.async.bulk + mbarrier ring and group equal keys in a shared open-addressing table, so a bucket of any size fits as
// long as its DISTINCT keys do -- a mitochondrial (k+1)-mer seen 8000 times is one slot instead of a bucket that no
// shared memory holds (the general kernel needs the whole bucket on chip and falls to chunk/merge/serial paths).
//
// A wide key does not fit a CAS, so a slot holds  fingerprint(40 bits) << 24 | representative(24 bits):  one 64-bit CAS
// claims the slot and publishes the bucket-relative index of the key that claimed it.  A later key whose fingerprint
// matches is compared word by word with that representative (a read of the bucket's own, just-streamed records: L2);
// unequal -> it is a different key and probing continues, so the result is exact whatever the fingerprints do.
#pragma once
#include "common.cuh"
#include "count_stream.cuh"
#include "local.cuh"

namespace mf {

constexpr int kCwNT = 512;
constexpr int kCwSlotsLog = 12;
constexpr int kCwSlots = 1 << kCwSlotsLog;
constexpr int kCwStages = 3;
constexpr int kCwWin = 512;
constexpr int kCwProbeLimit = 64;
constexpr int kCwPairsMax = 96;
template <int W>
struct CwCfg {
  // keys per ring stage: as many as two CTAs per SM allow (a thread's per-chunk bookkeeping is ~50 instructions, so one key
  // per thread and chunk doubles the cost of a key)
  static constexpr int kChunk = W == 3 ? 1024 : (W == 4 ? 768 : (W <= 6 ? 512 : 256));
  static constexpr int kSolidMax = W == 3 ? 768 : 512;     // distinct solid keys of one bucket on this path
  static constexpr int kAlign = (W % 4 == 0) ? 1 : ((W % 2 == 0) ? 2 : 4);   // records per 16-byte boundary
};
template <int W, int CH = CwCfg<W>::kChunk, int ST = kCwStages>
inline size_t count_stream_w_smem_bytes() {
  using C = CwCfg<W>;
  // slots u64[4096] | mbar u64[4] | ring u32[stages*chunk*W] (16-byte aligned) | skw u32[solid*W] | tcnt u32[4096]
  // | srep u32[solid] | scnt u32[solid] | bnd u32[win+2] | bins u32[1026] | small u32[64] | scratch u32[40] | flag i32[16]
  // | permA, permB, rk u16[solid]
  return (size_t)kCwSlots * 8 + 64 + (size_t)ST * CH * W * 4 + (size_t)C::kSolidMax * W * 4 + (size_t)kCwSlots * 4 +
         2 * (size_t)C::kSolidMax * 4 + (size_t)(kCwWin + 2) * 4 + 1026 * 4 + 64 * 4 + 40 * 4 + 16 * 4 + 3 * (size_t)C::kSolidMax * 2 + 16;
}

template <int W>
__device__ __forceinline__ void hash_wide(const uint32_t *k, uint32_t &ha, uint32_t &hb) {
  uint32_t a = 0x9e3779b9u, b = 0x7f4a7c15u;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    a = (a ^ k[i]) * 0x85ebca6bu;
    a ^= a >> 15;
    b = (b + k[i]) * 0xc2b2ae35u;
    b ^= b >> 13;
  }
  ha = a * 0x9e3779b1u;
  hb = b * 0x85ebca6bu;
}

// CH keys per ring stage, ST stages (the defaults, or 2 x 512 for W = 7, 8: every thread has a key in every chunk)
template <int W, int CH = CwCfg<W>::kChunk, int ST = kCwStages>
__global__ void __launch_bounds__(kCwNT) k_count_stream_w(LocalArgs a, const int32_t *__restrict__ cta_first) {
  extern __shared__ __align__(128) unsigned char smraw[];
  using C = CwCfg<W>;
  constexpr int NT = kCwNT, SMAX = C::kSolidMax, AL = C::kAlign;
  unsigned long long *slots = reinterpret_cast<unsigned long long *>(smraw);
  unsigned long long *mbar = slots + kCwSlots;
  uint32_t *ring = reinterpret_cast<uint32_t *>(mbar + 8);   // mbar[0..3] full, mbar[4..7] empty
  uint32_t *skw = ring + (size_t)ST * CH * W;
  uint32_t *tcnt = skw + (size_t)SMAX * W;
  uint32_t *srep = tcnt + kCwSlots;
  uint32_t *scnt = srep + SMAX;
  uint32_t *s_bnd = scnt + SMAX;
  uint32_t *bins = s_bnd + kCwWin + 2;
  uint32_t *s_small = bins + 1026;
  uint32_t *scratch = s_small + 64;
  if ((reinterpret_cast<uintptr_t>(scratch) & 7) != 0) scratch += 1;     // holds a u64 pair
  int *s_flag = reinterpret_cast<int *>(scratch + 40);   // 0 crowded, 1 ok, 2..3 arena base, 4 ns, 5 blk_left, 6..7 blk_pos
  uint16_t *permA = reinterpret_cast<uint16_t *>(s_flag + 16);
  uint16_t *permB = permA + SMAX, *rk = permB + SMAX;

  const int tid = threadIdx.x;
  const int b0 = cta_first[blockIdx.x], b1 = cta_first[blockIdx.x + 1];
  if (b0 >= b1) return;
  const int64_t rb = a.bkt_start[b0];
  const int64_t re = a.bkt_start[b1 - 1] + a.bkt_size[b1 - 1];
  const uint32_t total = (uint32_t)(re - rb);
  if (total == 0u) return;
  const int64_t A = rb - (rb % AL);                      // bulk copies need 16-byte aligned addresses and sizes
  const int64_t re_up = ((re + AL - 1) / AL) * AL;
  const int nchunks = (int)((re_up - A + CH - 1) / CH);
  const uint32_t m = (uint32_t)a.min_count;
  const int We = a.words_edge;

  auto issue = [&](int c) {   // thread 0 only
    const int s = c % ST;
    const int64_t g0 = A + (int64_t)c * CH;
    const int64_t left = re_up - g0;
    const uint32_t bytes = (uint32_t)(left < CH ? left : CH) * (uint32_t)(W * 4);
    mbar_expect_tx(mbar + s, bytes);
    bulk_g2s(ring + (size_t)s * CH * W, a.in + g0 * W, bytes, mbar + s);
  };
  int wb = b0;
  auto load_window = [&]() {
    for (int i = tid; i <= kCwWin; i += NT) {
      const int idx = wb + i;
      s_bnd[i] = idx < b1 ? (uint32_t)(a.bkt_start[idx] - rb) : total;
    }
  };
  if (tid == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(mbar + s, 1);              // full: the bulk copy's bytes
      mbar_init(mbar + 4 + s, NT / 32);    // empty: one arrival per warp
    }
    mbar_fence_init();
  }
  for (int i = tid; i < kCwSlots; i += NT) {
    slots[i] = kEmptyKey;
    tcnt[i] = 0u;
  }
  if (tid < 64) s_small[tid] = 0;
  if (tid < 16) s_flag[tid] = 0;
  load_window();
  __syncthreads();
  if (tid == 0)
    for (int c = 0; c < ST && c < nchunks; ++c) issue(c);

  uint32_t bbeg = 0;   // relative position of the current bucket's first key
  // key: W words in shared memory (the ring); rel: its position relative to the bucket start
  auto insert = [&](const uint32_t *key, uint32_t rel) {
    uint32_t ha, hb;
    hash_wide<W>(key, ha, hb);
    uint32_t h = ha >> (32 - kCwSlotsLog);
    const unsigned long long fp = (((unsigned long long)hb << 8) | (ha & 0xffu)) & 0xffffffffffull;
    const unsigned long long mine = (fp << 24) | rel;
    const uint32_t *bucket = a.in + (rb + (int64_t)bbeg) * W;
    for (int probe = 0; probe < kCwProbeLimit; ++probe) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(slots + h);
      if (cur == kEmptyKey) cur = atomicCAS(slots + h, kEmptyKey, mine);
      if (cur == kEmptyKey) {
        atomicAdd(tcnt + h, 1u);
        return;
      }
      if ((cur >> 24) == fp) {
        const uint32_t *rep = bucket + (size_t)(cur & 0xffffffull) * W;
        bool same = true;
#pragma unroll
        for (int j = 0; j < W; ++j) same = same && (__ldg(rep + j) == key[j]);
        if (same) {
          atomicAdd(tcnt + h, 1u);
          return;
        }
      }
      h = (h + 1) & (kCwSlots - 1);
    }
    s_flag[0] = 1;   // table too crowded: this bucket takes the general path
  };
  auto key_less = [&](const uint32_t *x, const uint32_t *y) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const uint32_t p2 = x[j], q2 = y[j];
      if (p2 != q2) return p2 < q2;
    }
    return false;
  };

  auto finish_bucket = [&](int slot) {
    const bool crowded = s_flag[0] != 0;
    for (int h = tid; h < kCwSlots; h += NT) {
      const uint32_t c = tcnt[h];
      if (c) {
        const uint32_t rep = (uint32_t)(slots[h] & 0xffffffull);
        slots[h] = kEmptyKey;
        tcnt[h] = 0u;
        if (a.counting && c < 64u) atomicAdd(s_small + c, 1u);
        if (c >= m) {
          const int q = atomicAdd(s_flag + 4, 1);
          if (q < SMAX) {
            srep[q] = rep;
            scnt[q] = c;
          }
        }
      } else if (slots[h] != kEmptyKey) {
        slots[h] = kEmptyKey;   // claimed by an insert that then gave up (crowded bucket)
      }
    }
    __syncthreads();
    const int ns_raw = s_flag[4];
    const bool bail = crowded || ns_raw > SMAX;
    const uint32_t ns = bail ? 0u : (uint32_t)ns_raw;
    if (a.counting) {
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
    // the solid keys themselves, from their representatives
    {
      const uint32_t *bucket = a.in + (rb + (int64_t)bbeg) * W;
      for (uint32_t x = tid; x < ns * W; x += NT) {
        const uint32_t q = x / W, j = x - q * W;
        skw[x] = __ldg(bucket + (size_t)srep[q] * W + j);
      }
    }
    if (tid == 0) {
      if (bail) {
        const int p = atomicAdd(a.bail_count, 1);
        a.bail_list[p] = slot;
        s_flag[1] = 0;
      } else if (ns > 0) {
        unsigned long long pos = ((unsigned long long)(uint32_t)s_flag[7] << 32) | (uint32_t)s_flag[6];
        int left = s_flag[5];
        if ((int)ns > left) {
          pos = atomicAdd(a.arena_cursor, (unsigned long long)kCsArenaBlock);
          left = kCsArenaBlock;
        }
        const int ok = pos + ns <= a.arena_cap;
        if (!ok) atomicExch(a.overflow_flag, 1);
        a.desc_off[slot] = (int64_t)pos;
        a.desc_cnt[slot] = ok ? (int64_t)ns : 0;
        s_flag[1] = ok;
        s_flag[2] = (int)(uint32_t)pos;
        s_flag[3] = (int)(uint32_t)(pos >> 32);
        pos += ns;
        left -= (int)ns;
        s_flag[5] = left;
        s_flag[6] = (int)(uint32_t)pos;
        s_flag[7] = (int)(uint32_t)(pos >> 32);
      } else {
        s_flag[1] = 0;
      }
    }
    if (ns > 1 && ns <= (uint32_t)kCwPairsMax)
      for (uint32_t q = tid; q < ns; q += NT) bins[q] = 0;
    __syncthreads();
    if (s_flag[1]) {
      const uint16_t *cur = nullptr;
      if (ns > 1 && ns <= (uint32_t)kCwPairsMax) {
        uint32_t nsp = 32;
        while (nsp < ns) nsp <<= 1;
        const uint32_t parts = NT / nsp, q = tid & (nsp - 1), part = tid / nsp;
        if (q < ns) {
          const uint32_t per = (ns + parts - 1) / parts;
          const uint32_t o0 = part * per, o1 = min(ns, o0 + per);
          const uint32_t *kq = skw + (size_t)q * W;
          uint32_t r = 0;
          for (uint32_t o = o0; o < o1; ++o) r += key_less(skw + (size_t)o * W, kq) ? 1u : 0u;
          if (r) atomicAdd(bins + q, r);
        }
        __syncthreads();
        for (uint32_t q2 = tid; q2 < ns; q2 += NT) permA[bins[q2]] = (uint16_t)q2;
        __syncthreads();
        cur = permA;
      } else if (ns > (uint32_t)kCwPairsMax) {
        // counting split on 10 bits of the 64-bit window that starts at the word holding the first unshared bit (all keys
        // of a bucket agree on every word before it, so the window orders them up to ties), then a rank fix per sub-bin
        const int j0 = min(a.bit_off >> 5, W - 1);
        auto window = [&](const uint32_t *k2) {
          return ((unsigned long long)k2[j0] << 32) | (j0 + 1 < W ? k2[j0 + 1] : 0u);
        };
        unsigned long long *s_mm = reinterpret_cast<unsigned long long *>(scratch);
        if (tid == 0) { s_mm[0] = ~0ull; s_mm[1] = 0ull; }
        for (int i = tid; i <= 1024; i += NT) bins[i] = 0u;
        __syncthreads();
        {
          unsigned long long mn = ~0ull, mx = 0ull;
          for (uint32_t q = tid; q < ns; q += NT) { const unsigned long long v = window(skw + (size_t)q * W); mn = min(mn, v); mx = max(mx, v); }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          }
          if ((tid & 31) == 0) { atomicMin(s_mm, mn); atomicMax(s_mm + 1, mx); }
        }
        __syncthreads();
        const unsigned long long vmin = s_mm[0];
        const int span_bits = 64 - __clzll((long long)((s_mm[1] - vmin) | 1ull));
        const int sh = span_bits > 10 ? span_bits - 10 : 0;
        for (uint32_t q = tid; q < ns; q += NT)
          rk[q] = (uint16_t)atomicAdd(bins + (uint32_t)((window(skw + (size_t)q * W) - vmin) >> sh), 1u);
        __syncthreads();
        block_excl_scan<NT>(bins, 1025, scratch + 4);
        for (uint32_t q = tid; q < ns; q += NT) permB[bins[(uint32_t)((window(skw + (size_t)q * W) - vmin) >> sh)] + rk[q]] = (uint16_t)q;
        __syncthreads();
        for (uint32_t p2 = tid; p2 < ns; p2 += NT) {
          const uint32_t q = permB[p2];
          const uint32_t *kq = skw + (size_t)q * W;
          const uint32_t d = (uint32_t)((window(kq) - vmin) >> sh);
          const uint32_t b2 = bins[d], e2 = bins[d + 1];
          uint32_t r = 0;
          for (uint32_t o = b2; o < e2; ++o) r += key_less(skw + (size_t)permB[o] * W, kq) ? 1u : 0u;
          permA[b2 + r] = (uint16_t)q;
        }
        __syncthreads();
        cur = permA;
      }
      const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
      for (uint32_t q = tid; q < ns; q += NT) {
        const uint32_t e = cur ? cur[q] : q;
        write_edge<W>(a.arena + (base + q) * (unsigned long long)We, skw + (size_t)e * W, We, scnt[e]);
      }
    }
    __syncthreads();
    if (tid == 0) { s_flag[0] = 0; s_flag[4] = 0; }
  };

  int cur_b = b0;
  uint32_t p = 0, bend = 0;
  auto advance = [&]() {
    while (cur_b < b1) {
      if (cur_b + 1 - wb > kCwWin) {
        __syncthreads();
        wb = cur_b;
        load_window();
        __syncthreads();
      }
      bbeg = s_bnd[cur_b - wb];
      bend = s_bnd[cur_b + 1 - wb];
      if (bend > p) break;
      ++cur_b;
    }
  };
  advance();
  const int off0 = (int)(rb - A);
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % ST;
    if constexpr (ST < 3) {   // two stages: the refill of the other stage cannot wait for the end of this chunk
      if (tid == 0 && c >= 1 && c - 1 + ST < nchunks) {
        mbar_wait(mbar + 4 + (c - 1) % ST, (uint32_t)(((c - 1) / ST) & 1));
        issue(c - 1 + ST);
      }
    }
    mbar_wait(mbar + s, (uint32_t)((c / ST) & 1));
    const int64_t g0 = A + (int64_t)c * CH;
    const int64_t ghi = g0 + CH < re ? g0 + CH : re;
    const uint32_t chi = (uint32_t)(ghi - rb);
    const uint32_t *rs = ring + (size_t)s * CH * W;
    const int off = off0 - c * CH;   // ring index of relative position q is q + off
    while (p < chi) {
      const uint32_t e = chi < bend ? chi : bend;
      const bool too_big = bend - bbeg > 0xffffffu;   // the representative index has 24 bits
      for (uint32_t q = p + tid; q < e; q += NT) {
        if (too_big) s_flag[0] = 1;
        else insert(rs + (size_t)((int)q + off) * W, q - bbeg);
      }
      p = e;
      if (p == bend) {
        __syncthreads();
        finish_bucket(cur_b);
        ++cur_b;
        advance();
        __syncthreads();
      }
    }
    // this warp is done with stage s; thread 0 refills the stage of the previous chunk once every warp has released it
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(mbar + 4 + s);
    if (ST >= 3 && tid == 0 && c >= 1 && c - 1 + ST < nchunks) {
      mbar_wait(mbar + 4 + (c - 1) % ST, (uint32_t)(((c - 1) / ST) & 1));
      issue(c - 1 + ST);
    }
  }
}

// distinct-ratio probe for W-word keys: the table holds a 64-bit hash of the key (the probe is a statistical estimate)
template <int W>
__global__ void k_probe_distinct_w(const uint32_t *__restrict__ keys, const ProbeChunk *__restrict__ chunks, int bit_off, uint32_t pattern,
                                   unsigned long long *tables, unsigned long long *stats) {
  const ProbeChunk ch = chunks[blockIdx.y];
  const uint32_t *src = keys + ch.start * W;
  unsigned long long *tab = tables + (size_t)ch.sample * kProbeSlots;
  const uint32_t want = ch.fbits ? (pattern & ((1u << ch.fbits) - 1u)) : 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ch.size; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[W];
#pragma unroll
    for (int j = 0; j < W; ++j) r[j] = src[i * W + j];
    if (ch.fbits && rec_digit<W>(r, bit_off, ch.fbits) != want) continue;
    uint32_t ha, hb;
    hash_wide<W>(r, ha, hb);
    unsigned long long key = ((unsigned long long)ha << 32) | hb;
    if (key == kEmptyKey) key = 0;
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

}  // namespace mf
