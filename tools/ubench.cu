// ubench.cu -- micro-benchmarks of the ranking primitives (peer mask of equal digits within a warp):
// match.any vs a ballot loop vs shared-memory atomics.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE, int NBITS>
__global__ void k(uint32_t *out, int iters, uint32_t seed) {
  __shared__ uint16_t sm[16 * 1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = seed ^ (blockIdx.x * 7919u + threadIdx.x * 104729u), acc = 0;
  uint16_t *wh = sm + warp * 1024;
  for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    const uint32_t d = (x >> 8) & ((1u << NBITS) - 1);
    if (MODE == 0) {
      unsigned m = __match_any_sync(0xffffffffu, d);
      acc += __popc(m) + __ffs(m);
    } else if (MODE == 1) {
      unsigned m = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < NBITS; ++b) {
        const unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1);
        m &= ((d >> b) & 1) ? v : ~v;
      }
      acc += __popc(m) + __ffs(m);
    } else if (MODE == 2) {
      acc += atomicAdd(reinterpret_cast<uint32_t *>(wh) + (d >> 1), 1u);
    } else if (MODE == 3) {   // full rank step with ballots: leader update + shfl
      unsigned m = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < NBITS; ++b) {
        const unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1);
        m &= ((d >> b) & 1) ? v : ~v;
      }
      const unsigned leader = __ffs(m) - 1;
      uint32_t old = 0;
      if (lane == leader) { old = wh[d]; wh[d] = old + __popc(m); }
      old = __shfl_sync(0xffffffffu, old, leader);
      acc += old + __popc(m & ((1u << lane) - 1));
      __syncwarp();
    } else if (MODE == 4) {   // same with match.any
      unsigned m = __match_any_sync(0xffffffffu, d);
      const unsigned leader = __ffs(m) - 1;
      uint32_t old = 0;
      if (lane == leader) { old = wh[d]; wh[d] = old + __popc(m); }
      old = __shfl_sync(0xffffffffu, old, leader);
      acc += old + __popc(m & ((1u << lane) - 1));
      __syncwarp();
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE, int NBITS>
void run(const char *name) {
  uint32_t *out;
  const int blocks = 148 * 4, threads = 512, iters = 2000;
  cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<MODE, NBITS><<<blocks, threads>>>(out, 10, 1);
  cudaEventRecord(a);
  k<MODE, NBITS><<<blocks, threads>>>(out, iters, 1);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  double keys = (double)blocks * threads * iters;
  printf("%-34s %8.3f ms  %8.1f Gkeys/s  (%.2f SM-cycles per warp-row at 1.965 GHz)\n", name, ms, keys / ms / 1e6,
         ms * 1e-3 * 1.965e9 * 148 / (keys / 32));
  cudaFree(out);
}
int main() {
  run<0, 10>("match.any 10b");
  run<1, 10>("ballot loop 10b");
  run<1, 8>("ballot loop 8b");
  run<1, 11>("ballot loop 11b");
  run<2, 10>("smem atomicAdd 10b (warp-private)");
  run<3, 10>("rank step, ballots 10b");
  run<3, 8>("rank step, ballots 8b");
  run<4, 10>("rank step, match.any 10b");
  return 0;
}
