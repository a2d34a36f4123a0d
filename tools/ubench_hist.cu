// ubench_hist.cu -- what bounds a streaming histogram on B200: the loads or the shared atomics?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_hist tools/ubench_hist.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_fill(uint4 *p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = i * 0x9e3779b97f4a7c15ull;
    x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 32;
    unsigned long long y = x * 0x94d049bb133111ebull; y ^= y >> 31;
    p[i] = make_uint4((uint32_t)x, (uint32_t)(x >> 32), (uint32_t)y, (uint32_t)(y >> 32));
  }
}
// MODE 0: loads + xor reduce; 1: + one shared atomic per 8-byte record on 1024 bins (top bits = random);
// 2: atomics on 64 lane-spread bins (bank-conflict free: bin*32+lane); 3: digit compute only (no atomic)
template <int MODE>
__global__ void __launch_bounds__(512) k_hist(const uint4 *__restrict__ p, size_t n, unsigned long long *out) {
  __shared__ uint32_t sh[2048 + 64];
  for (int i = threadIdx.x; i < 2048 + 64; i += 512) sh[i] = 0;
  __syncthreads();
  uint32_t acc = 0;
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t b = blockIdx.x * per, e = b + per < n ? b + per : n;
  for (size_t i0 = b; i0 < e; i0 += 512 * 8) {
    uint4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { size_t i = i0 + q * 512 + threadIdx.x; v[q] = p[i < e ? i : e - 1]; }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (MODE == 0) acc ^= v[q].x ^ v[q].z;
      if (MODE == 1) { atomicAdd(sh + (v[q].y >> 22), 1u); atomicAdd(sh + (v[q].w >> 22), 1u); }
      if (MODE == 2) { atomicAdd(sh + ((v[q].y >> 26) * 32 + (threadIdx.x & 31)), 1u); atomicAdd(sh + ((v[q].w >> 26) * 32 + (threadIdx.x & 31)), 1u); }
      if (MODE == 3) acc += (v[q].y >> 22) * 977u + (v[q].w >> 22);
    }
  }
  __syncthreads();
  if (MODE == 1 || MODE == 2) for (int i = threadIdx.x; i < 2048; i += 512) acc += sh[i];
  if (acc == 0x12345678u) out[0] = acc;
}
template <int MODE>
void run(const char *name, const uint4 *p, size_t n, unsigned long long *out, int ctas_per_sm) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) k_hist<MODE><<<148 * ctas_per_sm, 512>>>(p, n, out);
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) k_hist<MODE><<<148 * ctas_per_sm, 512>>>(p, n, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("%-34s ctas/SM %d: %7.3f ms  %6.0f GB/s  %6.1f G records/s\n", name, ctas_per_sm, ms, n * 16.0 / ms / 1e6, n * 2.0 / ms / 1e6);
}
int main() {
  const size_t n = (size_t)1 << 30;   // 16 GB of uint4 = 2^31 8-byte records
  uint4 *p; unsigned long long *out;
  cudaMalloc(&p, n * 16); cudaMalloc(&out, 64);
  k_fill<<<148 * 8, 512>>>(p, n);
  cudaDeviceSynchronize();
  for (int c : {2, 4}) {
    run<0>("loads only", p, n, out, c);
    run<3>("loads + digit", p, n, out, c);
    run<1>("loads + shared atomics (1024 bins)", p, n, out, c);
    run<2>("loads + conflict-free atomics", p, n, out, c);
  }
  // sustained: half a second of back-to-back launches, the last 50 timed (power / clock behaviour under load)
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 150; ++w) k_hist<1><<<148 * 4, 512>>>(p, n, out);
    cudaEventRecord(e0);
    for (int w = 0; w < 50; ++w) k_hist<1><<<148 * 4, 512>>>(p, n, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 50;
    printf("sustained shared-atomic histogram: %7.3f ms  %6.0f GB/s\n", ms, n * 16.0 / ms / 1e6);
  }
  // a 40 GB buffer (the bench's key buffers are 34 GB each)
  {
    cudaFree(p);
    const size_t n2 = (size_t)5 << 29;
    if (cudaMalloc(&p, n2 * 16) == cudaSuccess) {
      k_fill<<<148 * 8, 512>>>(p, n2);
      cudaDeviceSynchronize();
      run<1>("40 GB: loads + shared atomics", p, n2, out, 4);
      run<0>("40 GB: loads only", p, n2, out, 4);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
