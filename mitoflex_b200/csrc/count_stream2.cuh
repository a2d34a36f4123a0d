// count_stream2.cuh -- the streamed count finish for keys of 33..64 bits (16 <= k <= 31), second generation.
//
// Same contract as round 1's k_count_stream: persistent CTAs own contiguous runs of buckets, the keys arrive
// through a cp.async.bulk + mbarrier ring, a shared hash table groups them, the bucket end turns the solid keys into ordered
// edge records (KmerCounter::PackEdge).  What changed, and why (ncu r1i/r2a: the old kernel issued ~120 thread-instructions per
// key, 0.68 issue/cycle, stalls wait / barrier / branch_resolving):
//
//   * THE FIRST PROBE IS STRAIGHT-LINE CODE.  The old loop handed every lane one key and looped until the longest linear
//     probe of the warp was done (5-8 rounds at 45 % load although the AVERAGE probe is 1.6 slots).  Now a thread takes 4 keys
//     per round, has their 4 ring loads and 4 slot loads in flight together, settles what the first slot decides (a hit or a
//     claim: ~85 % at 22 % load) without a loop, and only the few keys left over probe on.  (A first attempt that handed out
//     keys to idle lanes with ballot + popc packed the lanes perfectly but spent ~150 instructions per warp iteration on the
//     bookkeeping: 12.5 ms against 8.0 ms of the old kernel on the 1.8 Gbp sample, ncu r2d.)
//   * ONE 64-BIT SLOT HOLDS KEY AND COUNT.  All keys of a bucket lie in a narrow range (level 1 fixed their top bits, the range
//     partition of level 2 their next ~10), so a slot stores the low RB bits of the right-aligned key above a CB-bit count
//     (RB + CB = 64): claiming is one CAS, counting one RED on the same word, the probe compares one shifted word.  The table
//     has 8192 slots in the memory the 4096 separate key / count slots took: half the load, ~1.15 probes per key.  A reference
//     key per bucket restores the full key (delta sign-extended in RB bits).  The host picks RB from the plan; when the range
//     does not fit (small inputs with few level-1 bits) the FULL variant keeps 64-bit keys and separate counts.
//   * NOBODY POLLS FOR A FREE STAGE.  A warp hands a ring stage back as soon as its keys of the chunk are in registers (a shared
//     counter per stage); the warp that hands it back last issues the bulk copy of the chunk after next itself.  (A dedicated
//     producer warp spinning on an "empty" mbarrier cost 8-9 % of all issued instructions, ncu r2h/r2i.)
#pragma once
#include "common.cuh"
#include "count_stream.cuh"

namespace mf {

constexpr int kC2NC = 512;              // consumer threads
constexpr int kC2NW = kC2NC / 32;       // consumer warps
constexpr int kC2NT = kC2NC;            // no producer warp: the last warp to hand a stage back issues its refill
constexpr int kC2RelSlotsLog = 13;      // REL: 8192 slots of 8 bytes (key | count)
constexpr int kC2FullSlotsLog = 12;     // FULL: 4096 keys + 4096 counts
constexpr int kC2ChunkLog = 11;         // 2048 keys (16 KB) per ring stage = one round of 4 keys per consumer thread
constexpr int kC2Chunk = 1 << kC2ChunkLog;
constexpr int kC2StagesRel = 2;         // ring stages (a power of two: the ring is addressed modulo stages * chunk).  Two are
constexpr int kC2StagesFull = 2;        // enough because a warp hands a stage back as soon as its keys are in registers
constexpr int kC2QueueCap = 128;        // per-warp queue (one round: 4 x 32 keys) of key parts the first slot did not settle; aliases the bucket-end arrays
constexpr int kC2MinCountBits = 17;     // REL needs room for counts up to 65535 and then some

template <bool REL>
struct C2Cfg {
  static constexpr int SlotsLog = REL ? kC2RelSlotsLog : kC2FullSlotsLog;
  static constexpr int Slots = 1 << SlotsLog;
  static constexpr int Stages = REL ? kC2StagesRel : kC2StagesFull;
  static constexpr int RingKeys = Stages * kC2Chunk;
  static constexpr size_t TableBytes = REL ? (size_t)Slots * 8 : (size_t)Slots * 12;
};
template <bool REL>
inline size_t count_stream2_smem_bytes() {
  using C = C2Cfg<REL>;
  // table | ring u64 | mbar u64[8] | bnd u32[win+2] | small u32[64] | flag i32[16] | ref u64[2]
  // | { skeys u64[768] | scnt u32[768] | bins u32[260] | scratch u32[40] | permA,permB,rk u16[768] }  (aliased by the queues)
  return C::TableBytes + (size_t)C::RingKeys * 8 + (size_t)kCsSolidMax * 8 + 64 + (size_t)kCsSolidMax * 4 + (size_t)(kCsWin + 2) * 4 +
         260 * 4 + 64 * 4 + 40 * 4 + 16 * 4 + 16 + 3 * (size_t)kCsSolidMax * 2;
}

__device__ __forceinline__ void c2_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kC2NC) : "memory"); }

// block_excl_scan for the 512 consumers (named barrier 1)
__device__ __forceinline__ uint32_t c2_excl_scan(uint32_t *s, int n, uint32_t *scratch) {
  constexpr int NT = kC2NC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + NT - 1) / NT;
  const int b = tid * per, e = min(n, b + per);
  uint32_t sum = 0;
  for (int i = b; i < e; ++i) sum += s[i];
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) scratch[warp] = inc;
  c2_sync();
  if (warp == 0) {
    uint32_t v = lane < NT / 32 ? scratch[lane] : 0;
    uint32_t vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
      if (lane >= o) vi += t;
    }
    scratch[lane] = vi - v;
    if (lane == 31) scratch[32] = vi;
  }
  c2_sync();
  uint32_t run = scratch[warp] + inc - sum;
  for (int i = b; i < e; ++i) {
    uint32_t v = s[i];
    s[i] = run;
    run += v;
  }
  uint32_t total = scratch[32];
  c2_sync();
  return total;
}

// key_shift = 64 - 2(k+1) (keys are left-aligned in 64 bits); REL: count_bits = CB, a slot is (rel << CB) | count with rel =
// the low 64 - CB bits of the right-aligned key; every bucket's key range must span < 2^(63 - CB) (the host guarantees it).
// CBT: compile-time count width (32: the slot's high word IS the key part -- one compare, no shifts), 0 = count_bits at run time.
template <bool REL, int CBT>
__global__ void __launch_bounds__(kC2NT, 2) k_count_stream2(LocalArgs a, const int32_t *__restrict__ cta_first, int key_shift, int count_bits) {
  extern __shared__ __align__(128) unsigned char smraw[];
  using C = C2Cfg<REL>;
  constexpr int NC = kC2NC, NW = kC2NW, Slots = C::Slots, Stages = C::Stages, RingKeys = C::RingKeys;
  unsigned long long *tab = reinterpret_cast<unsigned long long *>(smraw);                 // REL: [8192] slots; FULL: [4096] keys
  uint32_t *tcnt = reinterpret_cast<uint32_t *>(tab + (REL ? 0 : Slots));                  // FULL: [4096] counts
  unsigned long long *ring = reinterpret_cast<unsigned long long *>(smraw + C::TableBytes);
  unsigned long long *mbar = ring + RingKeys;                                              // [0..3] full, [4..7] empty
  uint32_t *s_bnd = reinterpret_cast<uint32_t *>(mbar + 8);
  uint32_t *s_small = s_bnd + kCsWin + 2;
  int *s_flag = reinterpret_cast<int *>(s_small + 64);   // 0 crowded, 1 ok, 2..3 arena base, 4 ns, 5 blk_left, 6..7 blk_pos
  unsigned long long *s_ref = reinterpret_cast<unsigned long long *>(s_flag + 16);         // [0] reference key of the bucket
  // the bucket-end arrays and the per-warp queues of the rounds are never live together (a barrier separates the phases)
  unsigned long long *skeys = s_ref + 2;
  uint32_t *scnt = reinterpret_cast<uint32_t *>(skeys + kCsSolidMax);
  uint32_t *bins = scnt + kCsSolidMax;
  uint32_t *scratch = bins + 260;
  uint16_t *permA = reinterpret_cast<uint16_t *>(scratch + 40);
  uint16_t *permB = permA + kCsSolidMax, *rk = permB + kCsSolidMax;
  uint32_t *wqueue = reinterpret_cast<uint32_t *>(skeys) + (threadIdx.x >> 5) * kC2QueueCap;   // [kC2QueueCap] key parts
  uint32_t *whist32 = reinterpret_cast<uint32_t *>(tab);   // sub-bin counters [1025] of the many-solid-keys sort alias the swept table

  const int tid = threadIdx.x, lane = tid & 31;
  const int b0 = cta_first[blockIdx.x], b1 = cta_first[blockIdx.x + 1];
  if (b0 >= b1) return;
  const int64_t rb = a.bkt_start[b0], re = a.bkt_start[b1 - 1] + a.bkt_size[b1 - 1];
  const uint32_t total = (uint32_t)(re - rb);
  if (total == 0u) return;
  const int64_t A = rb & ~(int64_t)1;                    // bulk copies need 16-byte aligned addresses
  const int64_t re_up = (re + 1) & ~(int64_t)1;
  const int nchunks = (int)((re_up - A + kC2Chunk - 1) >> kC2ChunkLog);
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(a.in);

  uint32_t *s_rel = reinterpret_cast<uint32_t *>(mbar + 4);   // [Stages] warps that have handed the stage back (current chunk)
  auto issue = [&](int c) {   // one thread: bulk copy of chunk c into its stage
    const int st = c & (Stages - 1);
    const int64_t g0 = A + ((int64_t)c << kC2ChunkLog);
    const int64_t left = re_up - g0;
    const uint32_t bytes = (uint32_t)(left < kC2Chunk ? left : kC2Chunk) * 8u;
    mbar_expect_tx(mbar + st, bytes);
    bulk_g2s(ring + (size_t)st * kC2Chunk, src + g0, bytes, mbar + st);
  };
  if (tid == 0) {
    for (int s = 0; s < Stages; ++s) {
      mbar_init(mbar + s, 1);          // full: the bulk copy's bytes
      s_rel[s] = 0u;
    }
    mbar_fence_init();
  }
  for (int i = tid; i < (int)(C::TableBytes / 8); i += kC2NT) tab[i] = REL ? 0ull : (i < Slots ? kEmptyKey : 0ull);
  if (tid < 64) s_small[tid] = 0;
  if (tid < 16) s_flag[tid] = 0;
  __syncthreads();
  if (tid == 0)
    for (int c = 0; c < Stages && c < nchunks; ++c) issue(c);
  // a warp is done reading stage s of chunk c (all lanes; the keys live in registers / the table now); the last of the NW warps
  // refills the stage.  The fence orders this warp's reads of the stage before the count the last warp acts on.
  auto release_stage = [&](int c) {
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const int st = c & (Stages - 1);
      if (atomicAdd(s_rel + st, 1u) == (uint32_t)(NW - 1)) {
        __threadfence_block();   // the other warps' reads of the stage (fenced before their increments) are behind us
        s_rel[st] = 0u;          // nobody touches the counter again before the refill issued here has landed and been consumed
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (c + Stages < nchunks) issue(c + Stages);
      }
    }
  };

  // ---------------------------------------------------------------- consumers
  const uint32_t m = (uint32_t)a.min_count;
  const int We = a.words_edge;
  const uint32_t off0 = (uint32_t)(rb - A);               // ring / chunk coordinates of relative position q: q + off0
  const int CB = CBT ? CBT : count_bits;
  const unsigned long long cmask = REL ? ((1ull << CB) - 1ull) : 0ull;
  const unsigned lt = lanemask_lt();

  int wb = b0;   // first bucket of the boundary window
  auto load_window = [&]() {
    for (int i = tid; i <= kCsWin; i += NC) {
      const int idx = wb + i;
      s_bnd[i] = idx < b1 ? (uint32_t)(a.bkt_start[idx] - rb) : total;
    }
  };
  load_window();
  c2_sync();

  // ---- bucket end: sweep + clear, solid keys -> ordered edge records.  All consumers, after a c2_sync.
  auto finish_bucket = [&](int slot) {
    const bool crowded = s_flag[0] != 0;
    if constexpr (REL) {
      const unsigned long long ref = s_ref[0] >> key_shift;   // right-aligned reference key
      if (!a.counting) {
        // no multiplicity histogram wanted (read2sdbg, seq2sdbg's count): only the SOLID slots need work -- about one slot in
        // twenty at assembly depths -- so the sweep first marks them (16 slots per thread, branch-free), then visits the marked
        // ones (a warp-aggregated reservation in the solid list), then clears its 16 slots.  The plain loop below ran its
        // occupied-slot body 16 times per warp and bucket whatever the table held: 60 % of the bucket end (ncu r2q).
        constexpr int SPT = Slots / NC;
        uint32_t solid = 0u;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
          const uint32_t c = (uint32_t)(tab[tid + j * NC] & cmask);
          solid |= (c >= m ? 1u : 0u) << j;
        }
        while (solid) {
          const int j = __ffs(solid) - 1;
          solid &= solid - 1u;
          const unsigned long long s = tab[tid + j * NC];
          const unsigned act = __activemask();
          const int leader = __ffs(act) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(s_flag + 4, __popc(act));
          base = __shfl_sync(act, base, leader);
          const int q = base + __popc(act & lt);
          if (q < kCsSolidMax) {
            const long long d = ((long long)(((s >> CB) - ref) << CB)) >> CB;
            skeys[q] = (unsigned long long)((long long)ref + d) << key_shift;
            scnt[q] = (uint32_t)(s & cmask);
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < SPT; ++j) tab[tid + j * NC] = 0ull;
      } else
      for (int h = tid; h < Slots; h += NC) {
        const unsigned long long s = tab[h];
        if (s) {
          tab[h] = 0ull;
          const uint32_t c = (uint32_t)(s & cmask);
          if (a.counting && c < 64u) atomicAdd(s_small + c, 1u);
          if (c >= m) {
            const int q = atomicAdd(s_flag + 4, 1);
            if (q < kCsSolidMax) {
              // full key = ref + (rel - ref) sign-extended in RB bits
              const long long d = ((long long)(((s >> CB) - ref) << CB)) >> CB;
              skeys[q] = (unsigned long long)((long long)ref + d) << key_shift;
              scnt[q] = c;
            }
          }
        }
      }
    } else {
      for (int h = tid; h < Slots; h += NC) {
        const uint32_t c = tcnt[h];
        if (c) {
          const unsigned long long key = tab[h];
          tab[h] = kEmptyKey;
          tcnt[h] = 0u;
          if (a.counting && c < 64u) atomicAdd(s_small + c, 1u);
          if (c >= m) {
            const int q = atomicAdd(s_flag + 4, 1);
            if (q < kCsSolidMax) {
              skeys[q] = key;
              scnt[q] = c;
            }
          }
        }
      }
    }
    c2_sync();
    const int ns_raw = s_flag[4];
    const bool bail = crowded || ns_raw > kCsSolidMax;
    const uint32_t ns = bail ? 0u : (uint32_t)ns_raw;
    if (a.counting) {
      // multiplicity histogram of the distinct keys (<prefix>.counting): small counts from shared memory, counts >= 64 are
      // solid keys (the host routes --min-count > 64 with a histogram request to the general kernel).  A bucket that
      // bails is counted by the path that takes it instead.
      if (tid < 64) {
        const uint32_t v = s_small[tid];
        s_small[tid] = 0;
        if (v && !bail) atomicAdd(a.counting + tid, (unsigned long long)v);
      }
      for (uint32_t q = tid; q < ns; q += NC) {
        const uint32_t c = scnt[q];
        if (c >= 64u) atomicAdd(a.counting + (c > (uint32_t)kMaxMul ? (uint32_t)kMaxMul : c), 1ull);
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
        if ((int)ns > left) {   // next arena block (what is left of the old one is abandoned)
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
    if (ns > 1 && ns <= (uint32_t)kCsPairsMax)
      for (uint32_t q = tid; q < ns; q += NC) bins[q] = 0;
    c2_sync();
    if (s_flag[1]) {
      const uint16_t *cur = nullptr;
      if (ns > 1 && ns <= (uint32_t)kCsPairsMax) {
        // keys are distinct: rank = number of smaller keys; the comparisons of one key are split over NC / nsp threads
        uint32_t nsp = 32;
        while (nsp < ns) nsp <<= 1;
        const uint32_t parts = NC / nsp, q = tid & (nsp - 1), part = tid / nsp;
        if (q < ns) {
          const uint32_t per = (ns + parts - 1) / parts;
          const uint32_t o0 = part * per, o1 = min(ns, o0 + per);
          const unsigned long long kq = skeys[q];
          uint32_t r = 0;
          for (uint32_t o = o0; o < o1; ++o) r += skeys[o] < kq;
          if (r) atomicAdd(bins + q, r);
        }
        c2_sync();
        for (uint32_t q2 = tid; q2 < ns; q2 += NC) permA[bins[q2]] = (uint16_t)q2;
        c2_sync();
        cur = permA;
      } else if (ns > (uint32_t)kCsPairsMax) {
        // many solid keys (moderate coverage): one counting split on the 10 bits below the keys' common range, then a rank
        // fix inside each sub-bin (the keys are distinct and spread evenly, so sub-bins hold about one key)
        unsigned long long *s_mm = reinterpret_cast<unsigned long long *>(scratch);   // [0] min, [1] max (8-byte aligned)
        if (tid == 0) { s_mm[0] = ~0ull; s_mm[1] = 0ull; }
        for (int i = tid; i <= 1024; i += NC) whist32[i] = 0u;
        c2_sync();
        {
          unsigned long long mn = ~0ull, mx = 0ull;
          for (uint32_t q = tid; q < ns; q += NC) { const unsigned long long kq = skeys[q]; mn = min(mn, kq); mx = max(mx, kq); }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          }
          if (lane == 0) { atomicMin(s_mm, mn); atomicMax(s_mm + 1, mx); }
        }
        c2_sync();
        const unsigned long long kmin = s_mm[0];
        const int span_bits = 64 - __clzll((long long)((s_mm[1] - kmin) | 1ull));
        const int sh = span_bits > 10 ? span_bits - 10 : 0;
        for (uint32_t q = tid; q < ns; q += NC) rk[q] = (uint16_t)atomicAdd(whist32 + (uint32_t)((skeys[q] - kmin) >> sh), 1u);
        c2_sync();
        c2_excl_scan(whist32, 1025, scratch);
        for (uint32_t q = tid; q < ns; q += NC) permB[whist32[(uint32_t)((skeys[q] - kmin) >> sh)] + rk[q]] = (uint16_t)q;
        c2_sync();
        for (uint32_t p2 = tid; p2 < ns; p2 += NC) {
          const uint32_t q = permB[p2];
          const unsigned long long kq = skeys[q];
          const uint32_t d = (uint32_t)((kq - kmin) >> sh);
          const uint32_t b2 = whist32[d], e2 = whist32[d + 1];
          uint32_t r = 0;
          for (uint32_t o = b2; o < e2; ++o) r += skeys[permB[o]] < kq;
          permA[b2 + r] = (uint16_t)q;
        }
        c2_sync();
        for (int i = tid; i <= 1024; i += NC) whist32[i] = REL ? 0u : 0xffffffffu;   // give the table back clean
        cur = permA;
      }
      const unsigned long long base = ((unsigned long long)(uint32_t)s_flag[3] << 32) | (uint32_t)s_flag[2];
      for (uint32_t q = tid; q < ns; q += NC) {
        const uint32_t e = cur ? cur[q] : q;
        const unsigned long long key = skeys[e];
        const uint32_t kw[2] = {(uint32_t)(key >> 32), (uint32_t)key};
        write_edge<2>(a.arena + (base + q) * (unsigned long long)We, kw, We, scnt[e]);
      }
    }
    c2_sync();
    if (tid == 0) {
      s_flag[0] = 0;
      s_flag[4] = 0;
    }
    // the barrier the caller issues next publishes the reset
  };

  uint32_t pbeg = 0;             // relative start of the current bucket (its first key is the REL reference)
  // ---- consumer loop: chunk by chunk in lockstep (every warp waits for the chunk, takes its share, hands the stage back)
  int cur_b = b0;
  uint32_t p = 0, bend = 0;      // p = relative position processed so far, bend = relative end of bucket cur_b
  auto advance = [&]() {         // first bucket at or after cur_b that ends beyond p (all consumers, uniform)
    while (cur_b < b1) {
      if (cur_b + 1 - wb > kCsWin) {
        c2_sync();
        wb = cur_b;
        load_window();
        c2_sync();
      }
      bend = s_bnd[cur_b + 1 - wb];
      if (bend > p) break;
      ++cur_b;
    }
  };
  advance();
  bool skip = false;             // the current bucket is not taken here (see below); set when a bucket starts
  auto bucket_begin = [&](uint32_t blen) {
    // REL: a count must stay inside its CB-bit field, so a bucket with 2^CB or more keys goes to the bail list untouched
    skip = REL && CB < 32 && blen >= (1u << CB);
    if (skip && tid == 0) s_flag[0] = 1;
  };
  bucket_begin(bend - p);

  constexpr int U = 4;           // keys per thread and round: all their loads are in flight together
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & (Stages - 1);
    mbar_wait(mbar + s, (uint32_t)((c / Stages) & 1));
    const uint32_t chi = min((uint32_t)(((c + 1) << kC2ChunkLog) - off0), total);   // relative end of the chunk
    bool released = false;   // this warp has handed stage s back already (warp-uniform)
    while (p < chi) {
      const uint32_t e = chi < bend ? chi : bend;
      if (!skip) {
        if (REL && p == pbeg && tid == 0) {   // the bucket's first key is its reference (it is in the ring: p lies in this chunk)
          const unsigned long long raw = ring[(pbeg + off0) & (uint32_t)(RingKeys - 1)];
          s_ref[0] = (raw << 32) | (raw >> 32);
        }
        if constexpr (REL && CBT == 32) {
          // slot = (key part << 32) | count: the high word is compared as it is, the low word counts.  A record is two
          // big-endian words (x = high half of the key), so the key part is one funnel shift.  Per round: 4 keys per thread,
          // their ring and slot loads in flight together; a hit is one predicated RED; everything else (a new key, a slot held
          // by another key) goes to the warp's queue with one ballot and is settled afterwards by all 32 lanes, one queued key
          // each -- no lane idles in a divergent probe loop, and the round itself is branch-free.
          const uint2 *ring2 = reinterpret_cast<const uint2 *>(ring);
          // the trip count is WARP-uniform (the round uses warp collectives): bounded by the warp's first position
          for (uint32_t qw = p + (uint32_t)(tid & ~31); qw < e; qw += U * NC) {
            const uint32_t q0 = qw + (uint32_t)lane;
            // a round lies inside one chunk = one contiguous ring stage: one address, constant offsets
            const uint2 *rp = ring2 + ((qw + off0) & (uint32_t)(RingKeys - 1)) + lane;
            uint32_t v[U], h[U];
            uint2 sl[U];
            bool pend[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              pend[u] = q0 + u * NC < e;
              uint2 w = make_uint2(0u, 0u);
              if (pend[u]) w = rp[u * NC];
              v[u] = __funnelshift_r(w.y, w.x, key_shift);
              h[u] = (v[u] * 0x9E3779B1u) >> (32 - C::SlotsLog);
            }
            // the last keys of the chunk are in registers: the stage can be refilled while they are processed
            if (e == chi && qw - (uint32_t)(tid & ~31) + U * NC >= e && !released) {
              release_stage(c);
              released = true;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {   // x = count, y = key part (volatile: the table changes under us)
              const unsigned long long t = *reinterpret_cast<volatile unsigned long long *>(tab + h[u]);
              sl[u] = make_uint2((uint32_t)t, (uint32_t)(t >> 32));
            }
            uint32_t nq = 0;   // queued keys of this round (warp-uniform)
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const bool hit = pend[u] && sl[u].y == v[u] && sl[u].x != 0u;   // an occupied slot has a count >= 1
              if (hit) atomicAdd(reinterpret_cast<uint32_t *>(tab + h[u]), 1u);
              const bool more = pend[u] && !hit;
              const unsigned bal = __ballot_sync(0xffffffffu, more);
              if (more) wqueue[nq + (uint32_t)__popc(bal & lt)] = v[u];
              nq += (uint32_t)__popc(bal);
            }
            __syncwarp();
            for (uint32_t i = lane; i < nq; i += 32) {
              const uint32_t vv = wqueue[i];
              uint32_t hh = (vv * 0x9E3779B1u) >> (32 - C::SlotsLog);
              const unsigned long long claim = ((unsigned long long)vv << 32) | 1ull;
#pragma unroll 1
              for (int probe = 0;; ++probe) {
                unsigned long long sv = *reinterpret_cast<volatile unsigned long long *>(tab + hh);
                if (sv == 0ull) sv = atomicCAS(tab + hh, 0ull, claim);   // empty: claim it with count 1
                if (sv == 0ull) break;
                if ((uint32_t)(sv >> 32) == vv) { atomicAdd(reinterpret_cast<uint32_t *>(tab + hh), 1u); break; }
                if (probe >= kCsProbeLimit) { s_flag[0] = 1; break; }    // table too crowded: the bucket bails
                hh = (hh + 1) & (Slots - 1);
              }
            }
            __syncwarp();   // the next round overwrites the queue
          }
        } else
        for (uint32_t q0 = p + tid; q0 < e; q0 += U * NC) {
          unsigned long long key[U];
          uint32_t h[U];
          bool pend[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t q = q0 + u * NC;
            pend[u] = q < e;
            // records are two big-endian u32 words: word 0 (the key's high half) is the low half of the 8-byte load
            const unsigned long long raw = ring[(min(q, e - 1) + off0) & (uint32_t)(RingKeys - 1)];
            key[u] = (raw << 32) | (raw >> 32);
            if constexpr (REL) h[u] = ((uint32_t)(key[u] >> key_shift) * 0x9E3779B1u) >> (32 - C::SlotsLog);
            else h[u] = ((((uint32_t)key[u] ^ ((uint32_t)(key[u] >> 32) * 0x9E3779B1u))) * 0x85EBCA6Bu) >> (32 - C::SlotsLog);
          }
          // first probe of every key; what it does not settle stays pending
          if constexpr (REL) {
            unsigned long long sl[U];
#pragma unroll
            for (int u = 0; u < U; ++u) sl[u] = *reinterpret_cast<volatile unsigned long long *>(tab + h[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (pend[u]) {
                const unsigned long long v = (key[u] >> key_shift) & (~0ull >> CB);
                unsigned long long sv = sl[u];
                if (sv == 0ull) sv = atomicCAS(tab + h[u], 0ull, (v << CB) | 1ull);
                if (sv == 0ull) {
                  pend[u] = false;
                } else if ((sv >> CB) == v) {
                  atomicAdd(reinterpret_cast<uint32_t *>(tab + h[u]), 1u);   // the count is the low word (CB <= 32 bits of it)
                  pend[u] = false;
                }
              }
            }
            // the rest: linear probing (a few lanes, a few steps)
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (pend[u]) {
                const unsigned long long v = (key[u] >> key_shift) & (~0ull >> CB);
                uint32_t hh = h[u];
                int probe = 1;
#pragma unroll 1
                for (;;) {
                  hh = (hh + 1) & (Slots - 1);
                  unsigned long long sv = *reinterpret_cast<volatile unsigned long long *>(tab + hh);
                  if (sv == 0ull) sv = atomicCAS(tab + hh, 0ull, (v << CB) | 1ull);
                  if (sv == 0ull) break;
                  if ((sv >> CB) == v) { atomicAdd(reinterpret_cast<uint32_t *>(tab + hh), 1u); break; }
                  if (++probe >= kCsProbeLimit) { s_flag[0] = 1; break; }   // table too crowded: the bucket bails
                }
              }
            }
          } else {
            unsigned long long sl[U];
#pragma unroll
            for (int u = 0; u < U; ++u) sl[u] = *reinterpret_cast<volatile unsigned long long *>(tab + h[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (pend[u]) {
                unsigned long long cur = sl[u];
                if (cur == kEmptyKey) cur = atomicCAS(tab + h[u], kEmptyKey, key[u]);
                if (cur == kEmptyKey || cur == key[u]) {
                  atomicAdd(tcnt + h[u], 1u);
                  pend[u] = false;
                }
              }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (pend[u]) {
                uint32_t hh = h[u];
                int probe = 1;
#pragma unroll 1
                for (;;) {
                  hh = (hh + 1) & (Slots - 1);
                  unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tab + hh);
                  if (cur == kEmptyKey) cur = atomicCAS(tab + hh, kEmptyKey, key[u]);
                  if (cur == kEmptyKey || cur == key[u]) { atomicAdd(tcnt + hh, 1u); break; }
                  if (++probe >= kCsProbeLimit) { s_flag[0] = 1; break; }
                }
              }
            }
          }
        }
      }
      p = e;
      if (p == bend) {
        c2_sync();
        finish_bucket(cur_b);
        ++cur_b;
        advance();
        pbeg = p;
        bucket_begin(bend - p);
        c2_sync();
      }
    }
    // this warp is done with stage s (its keys live in registers / the table now)
    if (!released) release_stage(c);
  }
}

}  // namespace mf
