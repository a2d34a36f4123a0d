// synth.cu -- synthetic PE reads generated directly in HBM (bench / test tooling, SURVEY.md 8d generator).
// Genome bases are a hash of their position (i.i.d. uniform ACGT), so nothing is stored: a 16.5 kb circular
// mitogenome sampled at `mito_fraction` of the pairs plus a linear nuclear background.  Per pair: insert ~
// N(mean, sd) clipped to [read_len, 600], uniform start, random strand, read 2 = reverse complement of the
// fragment end; substitution errors at `error_rate`; an N (rate `n_rate` per base, plus 0.1 % of reads with a
// leading and 0.1 % with a trailing run of 1-5 N) cuts the read the way megahit's FastxReader::TrimN does,
// because that is what reaches the packed library.
#include "engine.cuh"
#include "mfsdbg.h"

namespace mf {

__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t h3(uint64_t seed, uint64_t a, uint64_t b) {
  return mix64(seed ^ mix64(a * 0x632be59bd9b4e019ull + b * 0xd1342543de82ef95ull + 0x2545f4914f6cdd1dull));
}
__host__ __device__ inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

struct SynthP {
  int64_t n_pairs;
  int L;
  int64_t mito_len, nuc_len;
  double mito_fraction, error_rate, n_rate, ins_mean, ins_sd;
  uint64_t seed;
};
struct ReadGeom {   // one mate
  int is_mito;
  int64_t frag_start;
  int insert;
  int flip;      // fragment taken from the reverse strand
  int mate;      // 0/1
  int lead, len; // kept bases are [lead, lead+len) of the untrimmed mate
};
__device__ inline uint32_t genome_base(const SynthP &p, int is_mito, int64_t pos) {
  if (is_mito) {
    pos %= p.mito_len;
    if (pos < 0) pos += p.mito_len;
    return (uint32_t)(h3(p.seed, 0x4d49544full, (uint64_t)pos) & 3);
  }
  return (uint32_t)(h3(p.seed, 0x4e55434cull, (uint64_t)pos) & 3);
}
__device__ inline ReadGeom read_geom(const SynthP &p, int64_t read) {
  ReadGeom g;
  const uint64_t pair = (uint64_t)(read >> 1);
  g.mate = (int)(read & 1);
  g.is_mito = u01(h3(p.seed, pair, 1)) < p.mito_fraction;
  // Box-Muller
  double u1 = u01(h3(p.seed, pair, 2)), u2 = u01(h3(p.seed, pair, 3));
  if (u1 < 1e-300) u1 = 1e-300;
  double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  int ins = (int)llrint(p.ins_mean + p.ins_sd * z);
  ins = ins < p.L ? p.L : (ins > 600 ? 600 : ins);
  g.insert = ins;
  const int64_t G = g.is_mito ? p.mito_len : p.nuc_len;
  const int64_t span = g.is_mito ? G : (G - ins + 1 > 1 ? G - ins + 1 : 1);
  g.frag_start = (int64_t)(u01(h3(p.seed, pair, 4)) * (double)span);
  g.flip = (int)(h3(p.seed, pair, 5) & 1);
  // N model -> kept window
  int lead = 0, end = p.L;
  const uint64_t hr = h3(p.seed, (uint64_t)read, 6);
  if ((hr & 1023) == 0) lead = 1 + (int)((hr >> 10) % 5);
  if (((hr >> 20) & 1023) == 0) end = p.L - 1 - (int)((hr >> 30) % 5);
  // first interior N after `lead`: geometric with rate n_rate
  if (p.n_rate > 0) {
    double u = u01(h3(p.seed, (uint64_t)read, 7));
    if (u < 1e-300) u = 1e-300;
    double gap = floor(log(u) / log1p(-p.n_rate));
    if (gap < (double)(end - lead)) end = lead + (int)gap;
  }
  if (end < lead) end = lead;
  g.lead = lead;
  g.len = end - lead;
  return g;
}
// base `off` (0-based within the UNTRIMMED mate) of a read
__device__ inline uint32_t read_base(const SynthP &p, const ReadGeom &g, int64_t read, int off) {
  // position in fragment coordinates, fragment strand
  // mate 0 reads fragment[0..L) forward; mate 1 reads revcomp(fragment)[0..L)
  int fpos = g.mate ? g.insert - 1 - off : off;
  int comp = g.mate;
  // fragment -> genome
  int64_t gpos;
  if (g.flip) {
    gpos = g.frag_start + (g.insert - 1 - fpos);
    comp ^= 1;
  } else {
    gpos = g.frag_start + fpos;
  }
  uint32_t b = genome_base(p, g.is_mito, gpos);
  if (comp) b = 3 - b;
  const uint64_t he = h3(p.seed ^ 0x5eedull, (uint64_t)read, (uint64_t)off + 16);
  if (u01(he) < p.error_rate) b = (b + 1 + (uint32_t)((he & 0xff) % 3)) & 3;
  return b;
}

__global__ void k_synth_lengths(SynthP p, int64_t n_reads, int64_t *len) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_reads) len[r] = read_geom(p, r).len;
}
// three-kernel exclusive scan for int64 (n up to ~1e9): block sums, scan of sums (single block), fix-up
__global__ void k_block_sums(const int64_t *v, int64_t n, int64_t *sums) {
  __shared__ int64_t s[32];
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  int64_t x = i < n ? v[i] : 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    x = s[threadIdx.x];
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = x;
  }
}
__global__ void k_block_scan_fix(const int64_t *v, int64_t n, const int64_t *sum_off, int64_t *out) {
  __shared__ int64_t s[33];
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t x = i < n ? v[i] : 0, inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int64_t y = s[lane], yi = y;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, yi, o);
      if (lane >= o) yi += t;
    }
    s[lane] = yi - y;
  }
  __syncthreads();
  if (i < n) out[i] = sum_off[blockIdx.x] + s[warp] + inc - x;
  if (i == n - 1) out[n] = sum_off[blockIdx.x] + s[warp] + inc;
}
__global__ void k_scan_i64_1b(const int64_t *v, int64_t n, int64_t *out);   // defined below

void device_excl_scan_i64(Ctx &c, const int64_t *v, int64_t n, int64_t *out /* n+1 */) {
  if (n == 0) {
    MF_CUDA(cudaMemsetAsync(out, 0, 8, c.stream));
    return;
  }
  const int64_t nb = div_ceil64(n, 1024);
  int64_t *sums = nullptr;
  MF_CUDA(cudaMalloc(&sums, sizeof(int64_t) * (2 * nb + 2)));
  k_block_sums<<<(unsigned)nb, 1024, 0, c.stream>>>(v, n, sums);
  k_scan_i64_1b<<<1, 1024, 0, c.stream>>>(sums, nb, sums + nb + 1);
  k_block_scan_fix<<<(unsigned)nb, 1024, 0, c.stream>>>(v, n, sums + nb + 1, out);
  MF_LAUNCH_CHECK();
  c.launches += 3;
  MF_CUDA(cudaStreamSynchronize(c.stream));
  cudaFree(sums);
}
__global__ void k_scan_i64_1b(const int64_t *v, int64_t n, int64_t *out) {
  __shared__ int64_t s_part[1024];
  const int tid = threadIdx.x;
  const int64_t per = (n + blockDim.x - 1) / blockDim.x;
  const int64_t b = tid * per, e = b + per < n ? b + per : n;
  int64_t sum = 0;
  for (int64_t i = b; i < e; ++i) sum += v[i];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int64_t run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { int64_t x = s_part[i]; s_part[i] = run; run += x; }
    out[n] = run;
  }
  __syncthreads();
  int64_t run = s_part[tid];
  for (int64_t i = b; i < e; ++i) { int64_t x = v[i]; out[i] = run; run += x; }
}

// one thread per output word (16 bases)
__global__ void k_synth_words(SynthP p, int64_t n_reads, const int64_t *__restrict__ starts, int64_t n_bases,
                              uint32_t *__restrict__ words, int64_t n_words) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  int64_t g = w * 16;
  if (g >= n_bases) { words[w] = 0; return; }
  // read containing base g: last r with starts[r] <= g
  int64_t lo = 0, hi = n_reads;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (starts[mid] <= g) lo = mid; else hi = mid;
  }
  int64_t r = lo;
  ReadGeom geo = read_geom(p, r);
  int64_t rs = starts[r], re = starts[r + 1];
  uint32_t out = 0;
  for (int i = 0; i < 16 && g < n_bases; ++i, ++g) {
    while (g >= re) {   // skip to the read that owns g (empty reads possible)
      ++r;
      rs = starts[r];
      re = starts[r + 1];
      geo = read_geom(p, r);
    }
    out |= read_base(p, geo, r, geo.lead + (int)(g - rs)) << (30 - 2 * i);
  }
  words[w] = out;
}

void dev_synth(Ctx &c, const mfsdbg_synth_spec &sp, ReadsView *out) {
  if (sp.n_pairs < 0 || sp.read_len < 1 || sp.read_len > 600 || sp.mito_len < 1 || sp.nuclear_len < 601)
    throw std::invalid_argument("bad synthetic spec");
  SynthP p{sp.n_pairs, sp.read_len, sp.mito_len, sp.nuclear_len, sp.mito_fraction, sp.error_rate, sp.n_rate,
           sp.insert_mean, sp.insert_sd, sp.seed};
  const int64_t n_reads = 2 * sp.n_pairs;
  c.synth_starts.reserve(sizeof(int64_t) * (n_reads + 2));
  int64_t *d_starts = c.synth_starts.as<int64_t>();
  int64_t *d_len = nullptr;
  MF_CUDA(cudaMalloc(&d_len, sizeof(int64_t) * std::max<int64_t>(n_reads, 1)));
  if (n_reads > 0) {
    k_synth_lengths<<<(unsigned)div_ceil64(n_reads, 256), 256, 0, c.stream>>>(p, n_reads, d_len);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  device_excl_scan_i64(c, d_len, n_reads, d_starts);
  int64_t n_bases = 0;
  MF_CUDA(cudaMemcpyAsync(&n_bases, d_starts + n_reads, 8, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  cudaFree(d_len);
  const int64_t n_words = ((n_bases + 15) >> 4) + 16;   // 64 bytes of readable padding
  c.synth_words.reserve((size_t)n_words * 4);
  k_synth_words<<<(unsigned)div_ceil64(n_words, 256), 256, 0, c.stream>>>(p, n_reads, d_starts, n_bases, c.synth_words.as<uint32_t>(),
                                                                         n_words);
  MF_LAUNCH_CHECK();
  c.launches++;
  MF_CUDA(cudaStreamSynchronize(c.stream));
  out->packed = c.synth_words.as<uint32_t>();
  out->starts = d_starts;
  out->n_reads = n_reads;
  out->n_bases = n_bases;
}

}  // namespace mf
