// sdbg_ser.cuh -- the graph's records serialised on the device into the byte stream of `<prefix>.sdbg.N`
// (megahit SdbgWriter::Write: per item a 16-bit word  w | last<<4 | tip<<5 | min(mult, 255)<<8,  then a 16-bit multiplicity if
// mult > 254, then the tip's label words if it is a tip).  The host used to build that stream item by item (three vector
// inserts per item: 0.3 s for the 20 M items of BASELINE configs[0], most of `seq2sdbg`'s wall time next to 30 ms of kernels);
// now it only writes what arrives.
//
//   k_ser_tile_counts : per tile of 2048 items, how many are beyond 254 and how many are tips
//   k_ser_scan        : exclusive prefix of both over the tiles (one block; there are n / 2048 tiles)
//   k_ser_write       : a tile's items land at  tile_start + large_before + 2 wt tips_before  16-bit units; inside the tile a
//                       block scan of (units | tips << 16) per thread gives every thread its offset and its first tip label
#pragma once
#include "common.cuh"

namespace mf {

constexpr int kSerNT = 256, kSerIPT = 8, kSerTile = kSerNT * kSerIPT;

__global__ void __launch_bounds__(kSerNT) k_ser_tile_counts(const uint32_t *__restrict__ rec, int64_t n, uint32_t *__restrict__ tile_lt) {
  __shared__ uint32_t s_w[kSerNT / 32];
  const int64_t i0 = (int64_t)blockIdx.x * kSerTile + (int64_t)threadIdx.x * kSerIPT;
  uint32_t v = 0;
#pragma unroll
  for (int q = 0; q < kSerIPT; ++q) {
    if (i0 + q < n) {
      const uint32_t r = rec[i0 + q];
      v += ((r >> 8) > 254u ? 1u : 0u) + (((r >> 5) & 1u) << 16);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // <= 2048 per half: no carry
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kSerNT / 32; ++w) t += s_w[w];
    tile_lt[blockIdx.x] = t;
  }
}

// large_before[t], tips_before[t] for t = 0 .. ntiles (the last entry holds the totals)
__global__ void __launch_bounds__(1024) k_ser_scan(const uint32_t *__restrict__ tile_lt, int64_t ntiles, int64_t *__restrict__ large_before,
                                                   int64_t *__restrict__ tips_before) {
  __shared__ long long s_l[1024], s_t[1024];
  const int tid = threadIdx.x;
  const int64_t per = (ntiles + 1023) / 1024;
  const int64_t b = tid * per < ntiles ? tid * per : ntiles, e = b + per < ntiles ? b + per : ntiles;
  long long L = 0, T = 0;
  for (int64_t t = b; t < e; ++t) {
    const uint32_t v = tile_lt[t];
    L += v & 0xffffu;
    T += v >> 16;
  }
  s_l[tid] = L;
  s_t[tid] = T;
  __syncthreads();
  if (tid == 0) {
    long long al = 0, at = 0;
    for (int i = 0; i < 1024; ++i) {
      const long long l = s_l[i], t = s_t[i];
      s_l[i] = al;
      s_t[i] = at;
      al += l;
      at += t;
    }
    large_before[ntiles] = al;
    tips_before[ntiles] = at;
  }
  __syncthreads();
  L = s_l[tid];
  T = s_t[tid];
  for (int64_t t = b; t < e; ++t) {
    const uint32_t v = tile_lt[t];
    large_before[t] = L;
    tips_before[t] = T;
    L += v & 0xffffu;
    T += v >> 16;
  }
}

// tiles [tile0, tile0 + gridDim.x) into `out`, whose first 16-bit unit is unit `unit_base` of the whole stream
__global__ void __launch_bounds__(kSerNT) k_ser_write(const uint32_t *__restrict__ rec, const uint32_t *__restrict__ labels, int64_t n, int wt,
                                                      int64_t tile0, const int64_t *__restrict__ large_before,
                                                      const int64_t *__restrict__ tips_before, int64_t unit_base, uint16_t *__restrict__ out) {
  __shared__ uint32_t s[kSerNT];
  __shared__ uint32_t scratch[40];
  const int tid = threadIdx.x;
  const int64_t tile = tile0 + blockIdx.x;
  const int64_t i0 = tile * kSerTile + (int64_t)tid * kSerIPT;
  uint32_t r[kSerIPT];
  uint32_t mine = 0;
#pragma unroll
  for (int q = 0; q < kSerIPT; ++q) {
    r[q] = i0 + q < n ? rec[i0 + q] : 0u;
    if (i0 + q < n) {
      const uint32_t tip = (r[q] >> 5) & 1u;
      mine += 1u + ((r[q] >> 8) > 254u ? 1u : 0u) + tip * (uint32_t)(2 * wt) + (tip << 16);   // units < 65536 per tile: no carry
    }
  }
  s[tid] = mine;
  __syncthreads();
  block_excl_scan<kSerNT>(s, kSerNT, scratch);
  const uint32_t before = s[tid];
  int64_t u = tile * kSerTile + large_before[tile] + 2 * (int64_t)wt * tips_before[tile] - unit_base + (int64_t)(before & 0xffffu);
  int64_t t = tips_before[tile] + (int64_t)(before >> 16);
#pragma unroll
  for (int q = 0; q < kSerIPT; ++q) {
    if (i0 + q < n) {
      const uint32_t m = r[q] >> 8;
      out[u++] = (uint16_t)((r[q] & 0x3fu) | (min(m, 255u) << 8));
      if (m > 254u) out[u++] = (uint16_t)m;
      if (r[q] & 0x20u) {
        const uint32_t *lp = labels + t * wt;
        for (int w = 0; w < wt; ++w) {
          const uint32_t x = lp[w];
          out[u++] = (uint16_t)(x & 0xffffu);   // the label words as they lie in memory (little-endian)
          out[u++] = (uint16_t)(x >> 16);
        }
        ++t;
      }
    }
  }
}

}  // namespace mf
