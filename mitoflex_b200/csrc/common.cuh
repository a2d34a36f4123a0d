// common.cuh -- shared helpers for libmfsdbg (sm_100a only; no other arch is built).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

namespace mf {

constexpr int kNumBuckets = 65536;   // megahit bucket = first 8 bases (definitions.h kBucketPrefixLength)
constexpr int kMaxMul = 65535;       // mul_t = uint16
constexpr int kSentinel = 4;         // '$'
constexpr int kMaxDigitBits = 11;    // widest MSD digit a partition level uses
constexpr int kMaxBins = 1 << kMaxDigitBits;

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    throw CudaError(buf);
  }
}
#define MF_CUDA(x) ::mf::cuda_check((x), #x, __FILE__, __LINE__)
#define MF_LAUNCH_CHECK() ::mf::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

inline int div_ceil(int a, int b) { return (a + b - 1) / b; }
inline int64_t div_ceil64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// words of a canonical (k+1)-mer key, of an edge record, of an sdbg item, of a tip label
inline int words_key(int k) { return div_ceil(2 * (k + 1), 32); }
inline int words_edge(int k) { return div_ceil(2 * (k + 1) + 16, 32); }
inline int words_item(int k) { return div_ceil(2 * k + 4 + 16, 32); }
inline int words_tip(int k) { return div_ceil(2 * k, 32); }

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Compare two W-word big-endian records.
template <int W>
__device__ __forceinline__ int cmp_rec(const uint32_t *a, const uint32_t *b) {
#pragma unroll
  for (int i = 0; i < W; ++i) {
    uint32_t x = a[i], y = b[i];
    if (x != y) return x < y ? -1 : 1;
  }
  return 0;
}

// nbits (1..kMaxDigitBits) starting bit_off bits below the MSB of word 0 of a W-word record held in registers.
template <int W>
__device__ __forceinline__ uint32_t rec_digit(const uint32_t (&r)[W], int bit_off, int nbits) {
  if (bit_off + nbits <= 32) return (r[0] << bit_off) >> (32 - nbits);   // the partition levels live in word 0
  int wi = bit_off >> 5, sh = bit_off & 31;
  uint32_t hi = 0, lo = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    if (i == wi) hi = r[i];
    if (i == wi + 1) lo = r[i];
  }
  return __funnelshift_l(lo, hi, sh) >> (32 - nbits);
}
// same, record in memory
template <int W>
__device__ __forceinline__ uint32_t rec_digit_mem(const uint32_t *r, int bit_off, int nbits) {
  if (bit_off + nbits <= 32) return (r[0] << bit_off) >> (32 - nbits);
  int wi = bit_off >> 5, sh = bit_off & 31;
  uint32_t hi = r[wi], lo = (wi + 1 < W) ? r[wi + 1] : 0u;
  return __funnelshift_l(lo, hi, sh) >> (32 - nbits);
}

// Block-wide exclusive scan of s[0..n) in shared memory (uint32), returns total. All threads must call.
// scratch: at least 33 uint32 in shared memory.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t *s, int n, uint32_t *scratch) {
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
  __syncthreads();
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
  __syncthreads();
  uint32_t run = scratch[warp] + inc - sum;
  for (int i = b; i < e; ++i) {
    uint32_t v = s[i];
    s[i] = run;
    run += v;
  }
  uint32_t total = scratch[32];
  __syncthreads();
  return total;
}
#endif  // __CUDACC__

}  // namespace mf
