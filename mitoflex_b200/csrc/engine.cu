// engine.cu -- device pipelines of libmfsdbg: count, seq2sdbg, read2sdbg (sm_100a).
//
// count (megahit_core count, KmerCounter):
//   P0 k_reads_hist                   first key word of every position -> histogram of its top l1 bits       (reads.cuh)
//   P1 k_reads_scatter                keys recomputed and scattered by their top l1 bits; with LevelArgs::bin_base the
//                                     bins are peer-GPU buffers (fused partition + exchange)                   (reads.cuh)
//      k_probe_distinct               distinct / occurrences on a few prefix ranges -> bucket size      (count_stream.cuh)
//   P2 k_level_hist<RecordsProducer>  per-segment histogram of the range-partition digit                   (partition.cuh)
//   P3 k_scatter_tma / k_level_scatter  second partition level (TMA-fed tiles for 2-word records)
//   P4 k_count_stream / k_count_stream_w  persistent CTAs, bulk-copy ring, shared hash table, solid keys -> edge records;
//      k_count (local.cuh) for what bails and for 32-bit keys
//   P5 k_gather_edges                 compact per-bucket edge runs into the globally sorted edge array
// seq2sdbg (SeqToSdbg): item filter k_ks_* + k_items_real / k_items_miss (kmerset.cuh, k <= 63) / k_items_from_edges, k_items_from_seqs
//   (items.cuh), the same partition levels over item words, k_sdbg_local (sdbg_local.cuh; k_local<kSdbgEmit> for crowded
//   buckets) -> k_gather_edges.
#include "engine.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include "items.cuh"
#include "local.cuh"
#include "partition.cuh"
#include "reads.cuh"
#include "count_stream.cuh"
#include "count_stream2.cuh"
#include "kmerset.cuh"
#include "kmerset_w.cuh"
#include "count_stream_w.cuh"
#include "sdbg_local.cuh"

namespace mf {

// ------------------------------------------------------------------ memory
void DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return;
  // grow geometrically and in whole MiB: cudaMalloc / cudaFree stall for 100-450 ms at times (measured on these boxes), so a
  // buffer whose demand creeps up by a few per cent from call to call must not be reallocated every time
  size_t want = std::max(bytes, cap + cap / 2);
  want = (want + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
  release();
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {   // no room for the slack: take exactly what is needed
    cudaGetLastError();
    want = bytes;
    MF_CUDA(cudaMalloc(&p, want));
  }
  cap = want;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

void HostBuf::reserve(size_t bytes) {
  if (bytes <= cap) return;
  release();
  MF_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
  cap = bytes;
}
void HostBuf::release() {
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
}

Ctx::Ctx(int dev) : device(dev) {
  MF_CUDA(cudaSetDevice(dev));
  cudaDeviceProp prop;
  MF_CUDA(cudaGetDeviceProperties(&prop, dev));
  sm_count = prop.multiProcessorCount;
  MF_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
  stream = own_stream;
  edge_bucket_counts.assign(kNumBuckets, 0);
  sdbg_bucket_stats.assign((size_t)kNumBuckets * 3, 0);
}
Ctx::~Ctx() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  for (DevBuf *b : {&edges, &sdbg_rec, &sdbg_labels, &sdbg_buckets, &sbits, &pack_words, &pack_starts, &synth_words, &synth_starts,
                    &in_words, &in_starts})
    b->release();
  for (DevBuf &b : ov) b.release();
  miss.release();
  for (DevBuf &b : small) b.release();
  for (DevBuf &b : fb) b.release();
  out_rec.release();
  out_labels.release();
  out_large.release();
  for (HostBuf &b : io_pin) b.release();
  if (slab) cudaFree(slab);
  for (auto &s : stages) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
  for (auto e : copy_events) cudaEventDestroy(e);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (own_stream) cudaStreamDestroy(own_stream);
}
size_t Ctx::budget() {
  if (mem_limit) return mem_limit;
  size_t fr = 0, tot = 0;
  MF_CUDA(cudaMemGetInfo(&fr, &tot));
  mem_limit = (size_t)((double)(fr + slab_bytes) * 0.85);
  return mem_limit;
}
void Ctx::slab_reserve(size_t bytes) {
  slab_off = 0;
  if (bytes <= slab_bytes) return;
  MF_CUDA(cudaStreamSynchronize(stream));
  if (slab) MF_CUDA(cudaFree(slab));
  slab = nullptr;
  slab_bytes = 0;
  bytes = (bytes + (size_t)(64 << 20)) & ~(size_t)((1 << 20) - 1);
  cudaError_t e = cudaMalloc(&slab, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw CudaError("out of device memory reserving workspace of " + std::to_string(bytes >> 20) + " MiB");
  }
  slab_bytes = bytes;
}
void *Ctx::slab_alloc(size_t bytes) {
  size_t off = (slab_off + 255) & ~(size_t)255;
  if (off + bytes > slab_bytes) throw CudaError("workspace slab exhausted (internal sizing error)");
  slab_off = off + bytes;
  return slab + off;
}
void Ctx::d2h(void *dst, const void *src, size_t bytes) {
  if (bytes) MF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
  MF_CUDA(cudaStreamSynchronize(stream));
}
void Ctx::h2d(void *dst, const void *src, size_t bytes) {
  if (bytes) MF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
  MF_CUDA(cudaStreamSynchronize(stream));
}
static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void Ctx::begin_call() {
  call_t0 = now_ms();
  MF_CUDA(cudaSetDevice(device));
  for (auto &s : stages) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
  stages.clear();
  open_stages.clear();
  profile.clear();
}
void Ctx::end_call() {
  MF_CUDA(cudaStreamSynchronize(stream));
  if (profiling && getenv("MFSDBG_TRACE"))
    fprintf(stderr, "[mfsdbg] call took %.3f ms on the host clock (begin %.3f, end %.3f)\n", now_ms() - call_t0, fmod(call_t0, 1e6), fmod(now_ms(), 1e6));
  if (profiling) {
    char buf[128];
    for (auto &s : stages) {
      float ms = 0;
      cudaEventElapsedTime(&ms, s.e0, s.e1);
      snprintf(buf, sizeof buf, "%s=%.4f;", s.name.c_str(), ms);
      profile += buf;
      if (getenv("MFSDBG_TRACE")) fprintf(stderr, "[mfsdbg] %-20s host+%9.3f ms  device %9.3f ms\n", s.name.c_str(), s.host_ms, ms);
    }
  }
}
void Ctx::stage_begin(const char *name) {
  if (!profiling) return;
  StageRec r;
  r.name = name;
  r.host_ms = now_ms() - call_t0;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, stream);
  stages.push_back(r);
  open_stages.push_back((int)stages.size() - 1);
}
void Ctx::stage_end() {
  if (!profiling || open_stages.empty()) return;
  cudaEventRecord(stages[open_stages.back()].e1, stream);
  open_stages.pop_back();
}

#define MF_W_CASES(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10)
#define MF_DISPATCH_W(Wv, CALL)                                         \
  switch (Wv) {                                                         \
    MF_W_CASES(MF_DISPATCH_CASE_##CALL)                                 \
    default: throw std::runtime_error("unsupported record width (k too large, max k = 150)"); \
  }

template <class K>
static void set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) throw std::runtime_error("kernel shared memory request exceeds 227 KB");
  MF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// ------------------------------------------------------------------ small kernels
__global__ void k_start_bits(const int64_t *starts, int64_t n_reads, uint32_t *sbits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_reads) return;
  int64_t g = starts[i];
  atomicOr(sbits + (g >> 5), 1u << (g & 31));
}
__global__ void k_gather_edges(const uint32_t *__restrict__ arena, const int64_t *__restrict__ desc_off,
                               const int64_t *__restrict__ desc_cnt, const int64_t *__restrict__ out_off, int nslots, int We,
                               uint32_t *__restrict__ out) {
  int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= nslots) return;
  int64_t words = desc_cnt[slot] * We;
  const uint32_t *src = arena + desc_off[slot] * We;
  uint32_t *dst = out + out_off[slot] * We;
  for (int64_t x = threadIdx.x & 31; x < words; x += 32) dst[x] = src[x];
}
// edges are sorted: bucket b holds [lower_bound(b << 16), lower_bound((b + 1) << 16)) of the first key word
__global__ void k_edge_buckets(const uint32_t *edges, int64_t n, int We, unsigned long long *counts) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= kNumBuckets) return;
  auto lower = [&](uint32_t bucket) -> int64_t {   // first edge whose bucket >= `bucket`
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if ((edges[mid * We] >> 16) < bucket) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  counts[b] = (unsigned long long)((b + 1 < kNumBuckets ? lower((uint32_t)b + 1) : n) - lower((uint32_t)b));
}

// exclusive scan of int64 v[0..n) -> out[0..n] (out[n] = total): block sums, scan of the sums, fix-up (kernels in synth.cu);
// the scratch comes from the slab.  The single-block k_scan_i64 took 1.3 ms for the 524 288 bucket slots of the 5 Gbp sample.
__global__ void k_block_sums(const int64_t *v, int64_t n, int64_t *sums);
__global__ void k_block_scan_fix(const int64_t *v, int64_t n, const int64_t *sum_off, int64_t *out);
static void scan_slots(Ctx &c, const int64_t *v, int64_t n, int64_t *out) {
  if (n <= 4096) {
    k_scan_i64<<<1, 1024, 0, c.stream>>>(v, n, 0, out);
    MF_LAUNCH_CHECK();
    c.launches++;
    return;
  }
  const int64_t nb = div_ceil64(n, 1024);
  int64_t *sums = c.alloc<int64_t>((size_t)(2 * nb + 2));
  k_block_sums<<<(unsigned)nb, 1024, 0, c.stream>>>(v, n, sums);
  k_scan_i64<<<1, 1024, 0, c.stream>>>(sums, nb, 0, sums + nb + 1);
  k_block_scan_fix<<<(unsigned)nb, 1024, 0, c.stream>>>(v, n, sums + nb + 1, out);
  MF_LAUNCH_CHECK();
  c.launches += 3;
}

// ------------------------------------------------------------------ plan
struct Plan {
  int W, l1_bits, l2_bits, cap;
  int rest_bits;   // bits still to consume after level 1 (may need two more levels when a GPU owns few level-1 bins)
};
static int ceil_log2(double x) {
  int b = 0;
  while ((double)(1ull << b) < x && b < 62) ++b;
  return b;
}
static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
// part_limit: bits the partition levels may consume (count: 2(k+1); sdbg: 2(k-1), so that a (k-1)-prefix group never
// straddles a bucket).  density: how much denser than average the densest prefix range is (canonical keys = min of the two
// strands pile up at small prefixes with density 2(1-u); sdbg items come from both strands and are flat).
static Plan make_plan(int W, int part_limit, int64_t n_est, double density, bool hash_family, int forced_l1 = -1,
                      int nseg = 0) {
  Plan p;
  p.W = W;
  p.cap = hash_family ? local_cap(W, true, false) : sdbg_cap(W);
  // buckets of <= 64-bit keys are streamed through a key-resident table (any size, ~10-40 % of it distinct): ~5000 keys
  // on average keeps the densest ones (2x) well inside its 4096 slots; everything else must fit shared memory whole
  const double target = (hash_family && W <= 2) ? 5000.0 : p.cap * 0.65 / density;
  int bits = std::max(1, ceil_log2((double)std::max<int64_t>(n_est, 1) / target));
  bits = std::min(bits, 2 * kMaxDigitBits);
  bits = env_int(hash_family ? "MFSDBG_COUNT_BITS" : "MFSDBG_SDBG_BITS", bits);
  if (forced_l1 < 0) forced_l1 = env_int("MFSDBG_L1_BITS", -1);
  if (forced_l1 >= 0) {
    p.l1_bits = forced_l1;
  } else {
    p.l1_bits = bits <= 10 ? bits : bits / 2;   // the reads-fed level gets the narrower digit
  }
  p.l1_bits = std::max(1, std::min(p.l1_bits, std::min(kMaxDigitBits, part_limit)));
  p.l2_bits = std::max(0, std::min({bits - p.l1_bits, kMaxDigitBits, part_limit - p.l1_bits}));
  p.rest_bits = p.l2_bits;
  if (nseg > 0) {
    // the caller holds only `nseg` of the 2^l1 prefix bins (multi-GPU ownership): size the second level for those
    const int need = ceil_log2((double)std::max<int64_t>(n_est, 1) / ((double)nseg * target));
    p.rest_bits = std::max(0, std::min({need, 2 * kMaxDigitBits, part_limit - p.l1_bits}));
    p.l2_bits = std::min(p.rest_bits, kMaxDigitBits);
  }
  return p;
}

// ------------------------------------------------------------------ partition level launcher
template <int W>
struct TileCfg {
  static constexpr int NT = 512;
  static constexpr int IPT_S = W <= 2 ? 8 : (W <= 4 ? 4 : 2);   // scatter: records held in registers
  static constexpr int IPT_H = 16;                              // histogram: nothing held
  static constexpr int TS = NT * IPT_S;
  static constexpr int TH = NT * IPT_H;
};

// wide keys on short reads: few positions start a key (k = 141 on 150-base reads: 6 %, k = 119: 20 %), so the kernel that walks
// the LIST of valid positions (k_reads_scatter_compact, reads.cuh) wins over the one that walks all positions -- and, staging
// the keys of 8192 positions whatever their width, it takes 1024 level-1 bins where the walking kernel falls off a cliff beyond
// 256, which saves the whole prefix level that used to follow (count_l2a).  Measured on the 5 Gbp set, count at k = 141 / 119 /
// 99 / 79 / 59: 89.8 / 182.1 / 208.6 / 180.0 / 157.8 ms walking, 48.1 / 130.9 / 159.7 / 161.7 / 161.9 ms compacted with 1024
// bins (profiles/r2aa_wide_ab_matrix.json); with the later kernel (compact2, grouped copy-out) it also wins at k = 69 (192.6 -> 161.1 ms)
// and k = 59 (150.7 -> 143.6 ms; 61 % of the positions start a key), profiles/r2ag_wide_ab_matrix.json: taken below 62 %.
// MFSDBG_READS_COMPACT: 0 never, 1 by that rule, 2 always
static bool reads_compact_wanted(int W, int64_t n_bases, int64_t n_reads, int k) {
  if (W < 4) return false;
  const int mode = env_int("MFSDBG_READS_COMPACT", 1);
  if (mode == 0) return false;
  const double frac = n_bases > 0 ? std::max(0.0, (double)(n_bases - n_reads * (int64_t)k)) / (double)n_bases : 1.0;
  return mode == 2 || frac < 0.62;
}

// tile range [tile0, tile0 + ntiles) of the reads (ntiles < 0: all of them) -- the host-input pipeline runs the reads-fed
// level chunk by chunk while later chunks are still crossing PCIe
template <int W>
static void launch_reads_hist(Ctx &c, const ReadsView &r, const uint32_t *sbits, int k, LevelArgs a, unsigned long long *hist,
                              int64_t tile0 = 0, int64_t ntiles = -1, int64_t tstride = 1) {
  using C = ReadsTileCfg<W>;
  const int64_t tiles = ntiles >= 0 ? ntiles : div_ceil64(r.n_bases, C::T);
  if (tiles == 0) return;
  const size_t smem = (((size_t)1 << a.nbits) + 32 + reads_seq_words(C::NT, W) + reads_bit_words(C::NT, k)) * 4;
  const bool ranged = a.dlo != 0u || a.dhi != (1u << a.nbits);
  auto kern = ranged ? k_reads_hist<W, C::NT, true> : k_reads_hist<W, C::NT, false>;
  set_smem(kern, smem);
  const int64_t grid = std::min<int64_t>(tiles, (int64_t)c.sm_count * (2048 / C::NT));   // persistent: one flush per CTA
  kern<<<(unsigned)grid, C::NT, smem, c.stream>>>(ReadsSrc{r.packed, sbits, r.n_bases, k}, a, hist, tile0, tiles, tstride);
  MF_LAUNCH_CHECK();
  c.launches++;
}
template <int W, int BPT>
static void launch_reads_scatter_b(Ctx &c, const ReadsView &r, const uint32_t *sbits, int k, LevelArgs a, unsigned long long *cursor,
                                   uint32_t *out, int64_t tile0, int64_t tiles) {
  using C = ReadsTileCfg<W>;
  const size_t smem = reads_scatter_smem_bytes<W>(C::NT, a.nbits, k);
  const bool ranged = a.dlo != 0u || a.dhi != (1u << a.nbits);
  auto kern = ranged ? k_reads_scatter<W, C::NT, BPT, true> : k_reads_scatter<W, C::NT, BPT, false>;
  set_smem(kern, smem);
  kern<<<(unsigned)tiles, C::NT, smem, c.stream>>>(ReadsSrc{r.packed, sbits, r.n_bases, k}, a, cursor, out, tile0);
  MF_LAUNCH_CHECK();
  c.launches++;
}
template <int W>
static void launch_reads_scatter(Ctx &c, const ReadsView &r, const uint32_t *sbits, int k, LevelArgs a, unsigned long long *cursor,
                                 uint32_t *out, int64_t tile0 = 0, int64_t ntiles = -1) {
  using C = ReadsTileCfg<W>;
  const int64_t tiles = ntiles >= 0 ? ntiles : div_ceil64(r.n_bases, C::T);
  if (tiles == 0) return;
  if constexpr (W >= 4) {
    using CC = ReadsCompactCfg<W>;
    if ((1 << a.nbits) <= 2 * CC::NT && reads_compact_wanted(W, r.n_bases, r.n_reads, k)) {
      const int64_t pos0 = tile0 * C::T, pos1 = std::min<int64_t>(r.n_bases, pos0 + tiles * C::T);
      if (env_int("MFSDBG_READS_COMPACT_V", 2) == 2) {   // batches cut loose from the listing passes (reads.cuh); 1 = a batch per pass
        using C2 = ReadsCompact2Cfg<W>;
        const int64_t ctiles2 = div_ceil64(pos1 - pos0, C2::T);
        if (ctiles2 <= 0) return;
        const size_t smem2 = reads_compact2_smem_bytes<W>(a.nbits, k);
        auto kern2 = k_reads_scatter_compact2<W>;
        set_smem(kern2, smem2);
        kern2<<<(unsigned)ctiles2, C2::NT, smem2, c.stream>>>(ReadsSrc{r.packed, sbits, r.n_bases, k}, a, cursor, out, pos0, pos1);
        MF_LAUNCH_CHECK();
        c.launches++;
        return;
      }
      const int64_t ctiles = div_ceil64(pos1 - pos0, CC::T);
      if (ctiles <= 0) return;
      const size_t smem = reads_compact_smem_bytes<W>(a.nbits, k);
      auto kern = k_reads_scatter_compact<W>;
      set_smem(kern, smem);
      kern<<<(unsigned)ctiles, CC::NT, smem, c.stream>>>(ReadsSrc{r.packed, sbits, r.n_bases, k}, a, cursor, out, pos0, pos1);
      MF_LAUNCH_CHECK();
      c.launches++;
      return;
    }
  }
  const int bpt = std::max(1, (1 << a.nbits) / C::NT);   // bins per thread of the block scan
  switch (bpt) {
    case 1: launch_reads_scatter_b<W, 1>(c, r, sbits, k, a, cursor, out, tile0, tiles); break;
    case 2: launch_reads_scatter_b<W, 2>(c, r, sbits, k, a, cursor, out, tile0, tiles); break;
    case 4: launch_reads_scatter_b<W, 4>(c, r, sbits, k, a, cursor, out, tile0, tiles); break;
    case 8: launch_reads_scatter_b<W, 8>(c, r, sbits, k, a, cursor, out, tile0, tiles); break;
    default: launch_reads_scatter_b<W, 16>(c, r, sbits, k, a, cursor, out, tile0, tiles); break;
  }
}

// the records-fed scatter kernel for a digit width: bins per thread of its block scan is a template parameter
template <int W>
static auto level_scatter_kernel(int nbits) -> void (*)(RecordsProducer<W>, LevelArgs, unsigned long long *, uint32_t *) {
  using C = TileCfg<W>;
  const int bpt = std::max(1, (1 << nbits) / C::NT);
  if (bpt == 1) return k_level_scatter<RecordsProducer<W>, W, C::NT, C::IPT_S, 1>;
  if (bpt == 2) return k_level_scatter<RecordsProducer<W>, W, C::NT, C::IPT_S, 2>;
  return k_level_scatter<RecordsProducer<W>, W, C::NT, C::IPT_S, 4>;
}

struct HostChunks {
  std::vector<int64_t> start, size;
  std::vector<int32_t> seg;
  std::vector<int64_t> seg_out_start;   // [nseg]
  int nseg = 0;
};
struct DevBuckets {
  int64_t *start = nullptr, *size = nullptr;
  int nslots = 0;
};

// Partition `in` (chunks) by nbits at bit_off into `out`; segments land at seg_out_start.  Returns the bucket table
// (nseg << nbits slots, slot = seg * nbins + digit) allocated from `alloc`.
template <int W, class Alloc>
static DevBuckets partition_level(Ctx &c, const uint32_t *in, uint32_t *out, const HostChunks &hc, int bit_off, int nbits,
                                  Alloc &&alloc, const char *tag, const std::vector<uint16_t> *seg_nb = nullptr, int xbits = 0) {
  const std::string tag_h = std::string(tag) + "_hist", tag_s = std::string(tag) + "_scatter";
  using C = TileCfg<W>;
  const int nchunk = (int)hc.start.size(), nseg = hc.nseg, nbins = 1 << nbits;
  std::vector<int64_t> tb_h(nchunk + 1), tb_s(nchunk + 1);
  tb_h[0] = tb_s[0] = 0;
  for (int i = 0; i < nchunk; ++i) {
    tb_h[i + 1] = tb_h[i] + div_ceil64(hc.size[i], C::TH);
    tb_s[i + 1] = tb_s[i] + div_ceil64(hc.size[i], C::TS);
  }
  int64_t *d_start = (int64_t *)alloc(sizeof(int64_t) * nchunk), *d_size = (int64_t *)alloc(sizeof(int64_t) * nchunk);
  int32_t *d_seg = (int32_t *)alloc(sizeof(int32_t) * nchunk);
  int64_t *d_tbh = (int64_t *)alloc(sizeof(int64_t) * (nchunk + 1)), *d_tbs = (int64_t *)alloc(sizeof(int64_t) * (nchunk + 1));
  int64_t *d_sos = (int64_t *)alloc(sizeof(int64_t) * nseg);
  const size_t nsl = (size_t)nseg * nbins;
  unsigned long long *d_hist = (unsigned long long *)alloc(sizeof(unsigned long long) * nsl);
  unsigned long long *d_cur = (unsigned long long *)alloc(sizeof(unsigned long long) * nsl);
  DevBuckets b;
  b.start = (int64_t *)alloc(sizeof(int64_t) * nsl);
  b.size = (int64_t *)alloc(sizeof(int64_t) * nsl);
  b.nslots = (int)nsl;
  MF_CUDA(cudaMemcpyAsync(d_start, hc.start.data(), sizeof(int64_t) * nchunk, cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemcpyAsync(d_size, hc.size.data(), sizeof(int64_t) * nchunk, cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemcpyAsync(d_seg, hc.seg.data(), sizeof(int32_t) * nchunk, cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemcpyAsync(d_tbh, tb_h.data(), sizeof(int64_t) * (nchunk + 1), cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemcpyAsync(d_tbs, tb_s.data(), sizeof(int64_t) * (nchunk + 1), cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemcpyAsync(d_sos, hc.seg_out_start.data(), sizeof(int64_t) * nseg, cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nsl, c.stream));
  // the pageable host vectors above die with this frame: make sure the copies have been consumed
  MF_CUDA(cudaStreamSynchronize(c.stream));
  LevelArgs a{bit_off, nbits, 0u, (uint32_t)nbins, nullptr, nullptr, 0};
  if (seg_nb) {   // range partition: per-segment bin counts
    uint16_t *d_nb = (uint16_t *)alloc(sizeof(uint16_t) * nseg);
    c.h2d(d_nb, seg_nb->data(), sizeof(uint16_t) * nseg);
    a.seg_nb = d_nb;
    a.xbits = xbits;
  }
  TileDesc *d_tiles_h = (TileDesc *)alloc(sizeof(TileDesc) * std::max<int64_t>(tb_h[nchunk], 1));
  TileDesc *d_tiles_s = (TileDesc *)alloc(sizeof(TileDesc) * std::max<int64_t>(tb_s[nchunk], 1));
  if (tb_h[nchunk] > 0) {
    k_build_tiles<<<(unsigned)div_ceil64(tb_h[nchunk], 256), 256, 0, c.stream>>>(ChunkTable{d_start, d_size, d_seg, d_tbh, nchunk}, C::TH,
                                                                              tb_h[nchunk], d_tiles_h);
    k_build_tiles<<<(unsigned)div_ceil64(tb_s[nchunk], 256), 256, 0, c.stream>>>(ChunkTable{d_start, d_size, d_seg, d_tbs, nchunk}, C::TS,
                                                                              tb_s[nchunk], d_tiles_s);
    MF_LAUNCH_CHECK();
    c.launches += 2;
  }
  if (tb_h[nchunk] > 0) {
    const size_t smem = ((size_t)1 << nbits) * 4 + 16;
    Stage st(c, tag_h.c_str());
    if (bit_off < 32) {   // the 2-word fast path reads its digit with one funnel shift
      auto kern = k_level_hist_persist<W, C::NT>;
      set_smem(kern, smem);
      const int64_t grid = std::min<int64_t>(tb_h[nchunk], (int64_t)c.sm_count * 8);
      kern<<<(unsigned)grid, C::NT, smem, c.stream>>>(in, d_tiles_h, tb_h[nchunk], a, d_hist);
    } else {
      RecordsProducer<W> ph{in, d_tiles_h, C::TH};
      auto kern = k_level_hist<RecordsProducer<W>, W, C::NT, C::IPT_H>;
      set_smem(kern, smem);
      kern<<<(unsigned)tb_h[nchunk], C::NT, smem, c.stream>>>(ph, a, d_hist);
    }
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  k_level_scan<<<nseg, 1024, 0, c.stream>>>(d_hist, nbins, d_sos, d_cur, b.start, b.size);
  MF_LAUNCH_CHECK();
  c.launches++;
  if constexpr (W == 2) {
    if (tb_s[nchunk] > 0 && bit_off < 32) {
      // 2-word records: TMA-fed tiles of 6144 (5120 with 2048 bins) records
      // 512 threads x 12 keys (1024 threads x 6 keys at 32 registers measured the same: 18.8 against 18.3 ms, r2d)
      const int nt = 512;
      const int KPT = nbits <= 10 ? 12 : 10, T = nt * KPT;
      std::vector<int64_t> tb_t(nchunk + 1);
      tb_t[0] = 0;
      for (int i = 0; i < nchunk; ++i) tb_t[i + 1] = tb_t[i] + div_ceil64(hc.size[i], T);
      int64_t *d_tbt = (int64_t *)alloc(sizeof(int64_t) * (nchunk + 1));
      c.h2d(d_tbt, tb_t.data(), sizeof(int64_t) * (nchunk + 1));
      TileDesc *d_tiles_t = (TileDesc *)alloc(sizeof(TileDesc) * std::max<int64_t>(tb_t[nchunk], 1));
      k_build_tiles<<<(unsigned)div_ceil64(tb_t[nchunk], 256), 256, 0, c.stream>>>(ChunkTable{d_start, d_size, d_seg, d_tbt, nchunk}, T,
                                                                                tb_t[nchunk], d_tiles_t);
      MF_LAUNCH_CHECK();
      const int bpt = std::max(1, (1 << nbits) / nt);
      void (*kern)(const uint32_t *, const TileDesc *, int64_t, LevelArgs, unsigned long long *, uint32_t *) =
          nbits <= 10 ? (bpt == 1 ? k_scatter_tma<512, 12, 1> : k_scatter_tma<512, 12, 2>) : k_scatter_tma<512, 10, 4>;
      const size_t smem = nbits <= 10 ? scatter_tma_smem_bytes<12>(nbits) : scatter_tma_smem_bytes<10>(nbits);
      set_smem(kern, smem);
      Stage st(c, tag_s.c_str());
      const int64_t grid_t = std::min<int64_t>(tb_t[nchunk], (int64_t)c.sm_count * 2);
      kern<<<(unsigned)grid_t, nt, smem, c.stream>>>(in, d_tiles_t, tb_t[nchunk], a, d_cur, out);
      MF_LAUNCH_CHECK();
      c.launches += 2;
      return b;
    }
  }
  if constexpr (W >= 3 && W <= 5) {   // beyond 5 words a tile is under 2048 records and the register-staged kernel wins (measured at k=141)
    if (tb_s[nchunk] > 0 && bit_off < 32) {
      // wide records: the shared-memory staged kernel (tiles of 48 KB whatever the width)
      using SC = ScatterWCfg<W>;
      std::vector<int64_t> tb_t(nchunk + 1);
      tb_t[0] = 0;
      for (int i = 0; i < nchunk; ++i) tb_t[i + 1] = tb_t[i] + div_ceil64(hc.size[i], SC::T);
      int64_t *d_tbt = (int64_t *)alloc(sizeof(int64_t) * (nchunk + 1));
      c.h2d(d_tbt, tb_t.data(), sizeof(int64_t) * (nchunk + 1));
      TileDesc *d_tiles_t = (TileDesc *)alloc(sizeof(TileDesc) * std::max<int64_t>(tb_t[nchunk], 1));
      k_build_tiles<<<(unsigned)div_ceil64(tb_t[nchunk], 256), 256, 0, c.stream>>>(ChunkTable{d_start, d_size, d_seg, d_tbt, nchunk}, SC::T,
                                                                                tb_t[nchunk], d_tiles_t);
      MF_LAUNCH_CHECK();
      const int bpt = std::max(1, (1 << nbits) / 512);
      void (*kern)(const uint32_t *, const TileDesc *, int64_t, LevelArgs, unsigned long long *, uint32_t *) =
          bpt == 1 ? k_scatter_tma_w<W, 1> : (bpt == 2 ? k_scatter_tma_w<W, 2> : k_scatter_tma_w<W, 4>);
      const size_t smem = scatter_tma_w_smem_bytes<W>(nbits);
      set_smem(kern, smem);
      Stage st(c, tag_s.c_str());
      const int64_t grid_t = std::min<int64_t>(tb_t[nchunk], (int64_t)c.sm_count * 2);
      kern<<<(unsigned)grid_t, 512, smem, c.stream>>>(in, d_tiles_t, tb_t[nchunk], a, d_cur, out);
      MF_LAUNCH_CHECK();
      c.launches += 2;
      return b;
    }
  }
  if (tb_s[nchunk] > 0) {
    RecordsProducer<W> ps{in, d_tiles_s, C::TS};
    size_t smem = scatter_smem_bytes<W>(C::NT, C::TS, nbits, 4);
    auto kern = level_scatter_kernel<W>(nbits);
    set_smem(kern, smem);
    Stage st(c, tag_s.c_str());
    kern<<<(unsigned)tb_s[nchunk], C::NT, smem, c.stream>>>(ps, a, d_cur, out);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  return b;
}

// Consume `rest_bits` more key bits after level 1 with one or two partition levels.  `hc` describes the level-1 chunks
// (several may feed one segment); on return `*cur` holds the bucketed records and `*bit_off` the bits consumed.
template <int W, class Alloc>
static DevBuckets partition_chain(Ctx &c, uint32_t **cur, uint32_t **other, const HostChunks &hc, int *bit_off, int rest_bits,
                                  Alloc &&alloc, const char *tag) {
  DevBuckets b;
  const bool multi = (int)hc.start.size() != hc.nseg;
  if (rest_bits <= 0 && !multi) {   // level-1 bins are the buckets
    b.nslots = hc.nseg;
    b.start = (int64_t *)alloc(sizeof(int64_t) * b.nslots);
    b.size = (int64_t *)alloc(sizeof(int64_t) * b.nslots);
    c.h2d(b.start, hc.start.data(), sizeof(int64_t) * b.nslots);
    c.h2d(b.size, hc.size.data(), sizeof(int64_t) * b.nslots);
    return b;
  }
  rest_bits = std::max(rest_bits, 1);
  const int first = rest_bits <= kMaxDigitBits ? rest_bits : (rest_bits + 1) / 2;
  b = partition_level<W>(c, *cur, *other, hc, *bit_off, first, alloc, tag);
  std::swap(*cur, *other);
  *bit_off += first;
  const int second = rest_bits - first;
  if (second > 0) {
    std::vector<int64_t> st(b.nslots), sz(b.nslots);
    c.d2h(st.data(), b.start, sizeof(int64_t) * b.nslots);
    c.d2h(sz.data(), b.size, sizeof(int64_t) * b.nslots);
    HostChunks h2;
    h2.nseg = b.nslots;
    h2.start.assign(st.begin(), st.end());
    h2.size.assign(sz.begin(), sz.end());
    h2.seg.resize(b.nslots);
    for (int i = 0; i < b.nslots; ++i) h2.seg[i] = i;
    h2.seg_out_start = h2.start;
    const std::string tag3 = std::string(tag) + "b";
    b = partition_level<W>(c, *cur, *other, h2, *bit_off, second, alloc, tag3.c_str());
    std::swap(*cur, *other);
    *bit_off += second;
  }
  return b;
}

// ------------------------------------------------------------------ local finish launcher
template <int W, int MODE>
static void launch_local(Ctx &c, const LocalArgs &a, int grid) {
  if (grid <= 0) return;
  if constexpr (MODE == kSortOnly || MODE == kSdbgEmit) {
    size_t smem = local_smem_bytes(W, a.cap, false, false);
    auto kern = k_local<W, kLocalNT, MODE>;
    set_smem(kern, smem);
    kern<<<grid, kLocalNT, smem, c.stream>>>(a);
  } else {
    constexpr bool weighted = MODE == kCountMerge;
    size_t smem = local_smem_bytes(weighted ? W + 1 : W, a.cap, true, weighted);
    auto kern = k_count<W, kLocalNT, MODE>;
    set_smem(kern, smem);
    kern<<<grid, kLocalNT, smem, c.stream>>>(a);
  }
  MF_LAUNCH_CHECK();
  c.launches++;
}
template <int W>
static void launch_count_fast(Ctx &c, const LocalArgs &a, int grid) {
  if (grid <= 0) return;
  if constexpr (W <= 2) {
    size_t smem = fast_smem_bytes();
    auto kern = k_count_fast<W, kFastNT>;
    set_smem(kern, smem);
    kern<<<grid, kFastNT, smem, c.stream>>>(a);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
}
template <int W, int MODE>
static void launch_serial(Ctx &c, const LocalArgs &a, int nwork) {
  if (nwork <= 0) return;
  k_serial<W, MODE><<<div_ceil(nwork, 32), 32, 0, c.stream>>>(a, nwork);
  MF_LAUNCH_CHECK();
  c.launches++;
}

struct Range {
  int64_t start, size;
};
// Fully sort the given ranges of `cur` in place (all `sort_bits` leading bits matter); `other` is scratch with the same
// layout.  Recursive MSD levels; only reached by buckets with more DISTINCT keys than shared memory holds.
template <int W>
static void sort_ranges(Ctx &c, uint32_t *cur, uint32_t *other, const std::vector<Range> &ranges, int bit_off, int sort_bits, int depth = 0) {
  if (ranges.empty() || bit_off >= sort_bits) return;
  const int nbits = std::min(kMaxDigitBits, sort_bits - bit_off);
  // scratch from a grow-only buffer of the context (cudaMalloc / cudaFree stall for 100+ ms at times); recursion levels
  // take their own buffer so that a parent's tables survive
  int64_t total_size = 0;
  for (auto &r : ranges) total_size += r.size;
  const size_t need = (size_t)ranges.size() * ((size_t)kMaxBins * 40 + 256) + (size_t)(total_size / 128) + (2 << 20);
  DevBuf &scratch_buf = c.fb[std::min(depth, 3)];
  scratch_buf.reserve(need);
  size_t scratch_off = 0;
  auto alloc = [&](size_t bytes) {
    const size_t off = (scratch_off + 255) & ~(size_t)255;
    if (off + bytes > scratch_buf.cap) throw CudaError("fallback scratch exhausted (internal sizing error)");
    scratch_off = off + bytes;
    return (void *)((char *)scratch_buf.p + off);
  };
  HostChunks hc;
  hc.nseg = (int)ranges.size();
  for (int i = 0; i < hc.nseg; ++i) {
    hc.start.push_back(ranges[i].start);
    hc.size.push_back(ranges[i].size);
    hc.seg.push_back(i);
    hc.seg_out_start.push_back(ranges[i].start);
  }
  DevBuckets b = partition_level<W>(c, cur, other, hc, bit_off, nbits, alloc, "fallback");
  int32_t *d_bail = (int32_t *)alloc(sizeof(int32_t) * b.nslots);
  int *d_flags = (int *)alloc(sizeof(int) * 4);
  MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * 4, c.stream));
  LocalArgs a{};
  a.in = other;
  a.out = cur;
  a.bkt_start = b.start;
  a.bkt_size = b.size;
  a.work = nullptr;
  a.bit_off = bit_off + nbits;
  a.sort_bits = sort_bits;
  a.cap = local_cap(W, false, false);
  a.bail_list = d_bail;
  a.bail_count = d_flags;
  a.overflow_flag = d_flags + 1;
  {
    Stage st(c, "fallback_local");
    launch_local<W, kSortOnly>(c, a, b.nslots);
  }
  int nbail = 0;
  c.d2h(&nbail, d_flags, sizeof(int));
  if (getenv("MFSDBG_TRACE"))
    fprintf(stderr, "[mfsdbg] fallback sort depth %d: %d ranges, %lld records, %d slots, %d bail\n", depth, (int)ranges.size(),
            (long long)total_size, b.nslots, nbail);
  if (nbail > 0) {
    std::vector<int32_t> bl(nbail);
    std::vector<int64_t> st(b.nslots), sz(b.nslots);
    c.d2h(bl.data(), d_bail, sizeof(int32_t) * nbail);
    c.d2h(st.data(), b.start, sizeof(int64_t) * b.nslots);
    c.d2h(sz.data(), b.size, sizeof(int64_t) * b.nslots);
    std::vector<Range> sub;
    for (int s : bl) sub.push_back(Range{st[s], sz[s]});
    sort_ranges<W>(c, other, cur, sub, bit_off + nbits, sort_bits, depth + 1);
    for (auto &r : sub)
      MF_CUDA(cudaMemcpyAsync(cur + r.start * W, other + r.start * W, (size_t)r.size * W * 4, cudaMemcpyDeviceToDevice, c.stream));
  }
  MF_CUDA(cudaStreamSynchronize(c.stream));
}

// (start, size) of the listed slots, gathered on the device so that the whole table never crosses PCIe
__global__ void k_gather_slots(const int32_t *slots, int n, const int64_t *bkt_start, const int64_t *bkt_size, int64_t *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { out[2 * i] = bkt_start[slots[i]]; out[2 * i + 1] = bkt_size[slots[i]]; }
}
static std::vector<Range> fetch_bails(Ctx &c, const DevBuckets &b, const int32_t *d_bail, int nbail, std::vector<int32_t> *slots) {
  std::vector<Range> out;
  if (nbail <= 0) return out;
  slots->resize(nbail);
  c.d2h(slots->data(), d_bail, sizeof(int32_t) * nbail);
  std::sort(slots->begin(), slots->end());
  c.ov[8].reserve(sizeof(int32_t) * nbail);
  c.ov[9].reserve(sizeof(int64_t) * 2 * nbail);
  int32_t *d_s = c.ov[8].as<int32_t>();
  int64_t *d_o = c.ov[9].as<int64_t>();
  c.h2d(d_s, slots->data(), sizeof(int32_t) * nbail);
  k_gather_slots<<<div_ceil(nbail, 256), 256, 0, c.stream>>>(d_s, nbail, b.start, b.size, d_o);
  MF_LAUNCH_CHECK();
  c.launches++;
  std::vector<int64_t> ss(2 * (size_t)nbail);
  c.d2h(ss.data(), d_o, sizeof(int64_t) * 2 * nbail);
  for (int i = 0; i < nbail; ++i) out.push_back(Range{ss[2 * i], ss[2 * i + 1]});
  return out;
}

// ------------------------------------------------------------------ count (33..64-bit keys): probe + range partition
// distinct keys / key occurrences, measured on a few whole level-1 prefix ranges (see k_probe_distinct)
template <int W>
static double probe_distinct_ratio(Ctx &c, const uint32_t *keys, const HostChunks &l1, int bit_off, int key_bits) {
  constexpr int S = 8;
  std::vector<int64_t> seg_total(l1.nseg, 0);
  for (size_t i = 0; i < l1.start.size(); ++i) seg_total[l1.seg[i]] += l1.size[i];
  std::vector<int> nonempty;
  for (int s2 = 0; s2 < l1.nseg; ++s2)
    if (seg_total[s2] > 0) nonempty.push_back(s2);
  if (nonempty.empty()) return 1.0;
  std::vector<int> sample_of(l1.nseg, -1);
  // a sample is a whole segment scan: with the few, huge segments of a multi-GPU owner two of them are plenty
  int64_t seg_max = 1;
  for (int s2 : nonempty) seg_max = std::max(seg_max, seg_total[s2]);
  const int s_cap = (int)std::max<int64_t>(2, std::min<int64_t>(S, ((int64_t)256 << 20) / seg_max));
  const int ns = std::min<int>(s_cap, (int)nonempty.size());
  for (int j = 0; j < ns; ++j) sample_of[nonempty[(size_t)((2 * j + 1) * nonempty.size() / (2 * ns))]] = j;
  std::vector<ProbeChunk> pcs;
  for (size_t i = 0; i < l1.start.size(); ++i) {
    const int sm = sample_of[l1.seg[i]];
    if (sm < 0 || l1.size[i] == 0) continue;
    int fb = 0;
    while (fb < std::min(16, key_bits - bit_off - 8) && (seg_total[l1.seg[i]] >> (fb + 1)) >= 3000) ++fb;
    pcs.push_back(ProbeChunk{l1.start[i], l1.size[i], sm, fb});
  }
  unsigned long long *d_tab = c.alloc<unsigned long long>((size_t)S * kProbeSlots);
  unsigned long long *d_stats = c.alloc<unsigned long long>(2 * S);
  ProbeChunk *d_pcs = c.alloc<ProbeChunk>(pcs.size());
  MF_CUDA(cudaMemsetAsync(d_tab, 0xff, sizeof(unsigned long long) * S * kProbeSlots, c.stream));
  MF_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(unsigned long long) * 2 * S, c.stream));
  c.h2d(d_pcs, pcs.data(), sizeof(ProbeChunk) * pcs.size());
  {
    Stage st(c, "probe");
    // enough CTAs to stream the sampled segments at full bandwidth whatever the number of chunks
    const unsigned probe_gx = (unsigned)std::max<size_t>(16, std::min<size_t>(1024, (size_t)8 * c.sm_count / std::max<size_t>(pcs.size(), 1)));
    if constexpr (W == 2) k_probe_distinct<<<dim3(probe_gx, (unsigned)pcs.size()), 256, 0, c.stream>>>(keys, d_pcs, bit_off, 0x5a5au, d_tab, d_stats);
    else k_probe_distinct_w<W><<<dim3(probe_gx, (unsigned)pcs.size()), 256, 0, c.stream>>>(keys, d_pcs, bit_off, 0x5a5au, d_tab, d_stats);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  unsigned long long st[2 * S];
  c.d2h(st, d_stats, sizeof st);
  double occ = 0, dis = 0;
  std::vector<double> ratios;
  for (int j = 0; j < S; ++j) {
    occ += (double)st[2 * j];
    dis += (double)st[2 * j + 1];
    if (st[2 * j] >= 512) ratios.push_back((double)st[2 * j + 1] / (double)st[2 * j]);
  }
  if (occ < 64) return 1.0;
  // The MEDIAN of the samples' ratios, not the pooled ratio: a sample is a few thousand keys, and one high-multiplicity key in
  // it (a mitogenome (k+1)-mer: 13 000 copies) drags the pooled ratio down several-fold -- the buckets then come out that
  // much too large and a quarter of them overflow their tables (70 ms in the multi-pass kernel on one rank of eight, r2u).
  // Hot keys do not crowd a table; the typical bucket is what the size must fit.
  if (ratios.size() >= 3) {
    std::sort(ratios.begin(), ratios.end());
    const double med = ratios.size() % 2 ? ratios[ratios.size() / 2] : 0.5 * (ratios[ratios.size() / 2 - 1] + ratios[ratios.size() / 2]);
    if (getenv("MFSDBG_TRACE")) fprintf(stderr, "[mfsdbg] probe: pooled %.4f, median of %d samples %.4f\n", dis / occ, (int)ratios.size(), med);
    return std::max(med, 1e-4);
  }
  return std::max(dis / occ, 1e-4);
}

// Level 2 of the streamed count: buckets sized for the shared table from the measured distinct ratio; every level-1
// segment is cut into its own number of equal key ranges.  *bit_off keeps the bits ALL keys of a bucket share.
template <int W, class Alloc>
static DevBuckets stream_partition(Ctx &c, uint32_t **cur, uint32_t **other, const HostChunks &l1, int *bit_off, int key_bits,
                                   int min_count, double *rho_out, int *rel_count_bits, Alloc &&alloc) {
  const double rho = probe_distinct_ratio<W>(c, *cur, l1, *bit_off, key_bits);
  *rho_out = rho;
  // keys per bucket for a table of `slots` slots: 2-word keys whose slots hold "key | count" get 8192 of them (k_count_stream2<REL>),
  // the full-key variant and the wide-key kernel 4096
  const bool rel_wanted = W == 2 && env_int("MFSDBG_COUNT_REL", 1) != 0;
  const double load = env_int("MFSDBG_STREAM_LOAD_PCT", W == 2 ? (rel_wanted ? 42 : 45) : 35) / 100.0;
  const int solid_max = W == 2 ? kCsSolidMax : (W == 3 ? 768 : 512);
  auto bucket_keys = [&](int slots) {
    double v = std::min(60000.0 * slots / kCsSlots, std::max(1024.0, load * slots / rho));
    // --min-count 1 makes every distinct key a solid one: keep them under the per-bucket limit of the streamed kernel
    if (min_count <= 1 && load <= 1.0) v = std::min(v, std::max(512.0, 0.7 * solid_max / rho));
    return v;
  };
  bool b_is_rel = rel_wanted;
  double B = bucket_keys(rel_wanted ? C2Cfg<true>::Slots : kCsSlots);
  *rel_count_bits = 0;
  if (getenv("MFSDBG_TRACE")) fprintf(stderr, "[mfsdbg] distinct ratio %.4f -> %.0f keys per bucket\n", rho, B);
  HostChunks hc = l1;
  DevBuckets b;
  for (;;) {
    std::vector<int64_t> seg_total(hc.nseg, 0);
    for (size_t i = 0; i < hc.start.size(); ++i) seg_total[hc.seg[i]] += hc.size[i];
    const int xbits = std::min(16, key_bits - *bit_off);
    // 1024 bins keep the scatter's runs long enough (2048-bin levels run at half the speed); more only as a last resort
    int64_t nb_cap = std::min<int64_t>(1024, (int64_t)1 << xbits);
    int64_t max_need = 1;
    for (int s2 = 0; s2 < hc.nseg; ++s2) max_need = std::max<int64_t>(max_need, (int64_t)std::ceil((double)seg_total[s2] / B));
    // wide keys: a 2048-bin level runs 1.35x slower than a 1024-bin one (k=79: 49 against 36 ms), a whole extra prefix level
    // costs 34 ms -- take the 2048 bins when they make that level unnecessary (MFSDBG_L2_NBCAP_W=1024 turns it off)
    if (W >= 4 && xbits == 16 && max_need > nb_cap && max_need <= std::min(kMaxBins, env_int("MFSDBG_L2_NBCAP_W", 2048))) nb_cap = max_need;
    // up to 1.5x over the bin limit the buckets simply get that much larger (the table runs fuller, what crowds it takes the
    // multi-pass kernel): cheaper than a whole extra partition level
    // (2-word keys only: wider keys have no multi-pass kernel, their crowded buckets fall to the slow general path)
    if (max_need <= nb_cap + (W == 2 ? nb_cap / 2 : 0) || xbits < 16) {
      // 2-word keys: can a slot of the streamed kernel hold "low RB bits of the key | CB-bit count" (k_count_stream2<REL>)?  The
      // keys of a bucket span at most 2^(key_bits - bit_off) / nb of the key space; RB must cover twice that, CB >= 17.
      int64_t nb_min = 1;
      if (W == 2 && env_int("MFSDBG_COUNT_REL", 1) != 0) {
        const int rest = key_bits - *bit_off;                       // bits the keys of a segment may differ in
        int64_t tot = 0;
        for (int64_t v : seg_total) tot += v;
        const double avg_need = (double)tot / ((double)std::max(hc.nseg, 1) * B);
        const bool force = env_int("MFSDBG_COUNT_REL", 1) == 2;
        // cutting every segment into >= 2^cut ranges is free when the segments get about that many buckets anyway (or so few
        // that the empty ones do not matter)
        auto cut_ok = [&](int cut) {
          return cut <= 0 || (cut <= 10 && xbits == 16 && ((int64_t)1 << cut) <= nb_cap && (cut <= 4 || force || avg_need >= (double)((int64_t)1 << cut) * 0.5));
        };
        const int cut32 = rest - 31;                                // 32-bit key part + 32-bit count: the cheapest slot compare
        // (a variant with a run-time split of the 64 bits -- 47-bit key part + 17-bit count for k up to 31 -- worked but its
        // 64-bit shifts and compares made the round twice as long: 40.9 ms against 20.9 ms at k=29 / k=21 on the 5 Gbp sample;
        // those k take the full-key variant)
        if (cut_ok(cut32)) {
          nb_min = (int64_t)1 << std::max(cut32, 0);
          *rel_count_bits = 32;
        }
      }
      if (b_is_rel && *rel_count_bits == 0) {   // the buckets were sized for the 8192-slot table: size them for the full-key one
        b_is_rel = false;
        B = bucket_keys(kCsSlots);
        continue;
      }
      std::vector<uint16_t> nb(hc.nseg);
      int mx = 1;
      for (int s2 = 0; s2 < hc.nseg; ++s2) {
        nb[s2] = (uint16_t)std::max<int64_t>(nb_min, std::min<int64_t>(nb_cap, (int64_t)std::ceil((double)seg_total[s2] / B)));
        mx = std::max<int>(mx, nb[s2]);
      }
      const int nbits = std::max(1, ceil_log2((double)mx));
      b = partition_level<W>(c, *cur, *other, hc, *bit_off, nbits, alloc, "count_l2", &nb, xbits);
      std::swap(*cur, *other);
      return b;
    }
    // a segment needs more than 2048 buckets (shallow data / few segments): one bit-prefix level first
    const int bits = std::min(kMaxDigitBits, ceil_log2((double)max_need / (double)nb_cap));
    b = partition_level<W>(c, *cur, *other, hc, *bit_off, bits, alloc, "count_l2a");
    std::swap(*cur, *other);
    *bit_off += bits;
    std::vector<int64_t> st(b.nslots), sz(b.nslots);
    c.d2h(st.data(), b.start, sizeof(int64_t) * b.nslots);
    c.d2h(sz.data(), b.size, sizeof(int64_t) * b.nslots);
    HostChunks h2;
    h2.nseg = b.nslots;
    h2.start.assign(st.begin(), st.end());
    h2.size.assign(sz.begin(), sz.end());
    h2.seg.resize(b.nslots);
    for (int i = 0; i < b.nslots; ++i) h2.seg[i] = i;
    h2.seg_out_start = h2.start;
    hc = std::move(h2);
  }
}

// rel_count_bits > 0: slots hold "key low bits | count" (stream_partition guarantees the buckets' key ranges fit)
static void launch_count_stream(Ctx &c, const LocalArgs &a, int nslots, int64_t n_keys, int grid, int32_t *d_cta_first, int key_bits,
                                int rel_count_bits) {
  k_split_ranges<<<div_ceil(grid + 1, 128), 128, 0, c.stream>>>(a.bkt_start, a.bkt_size, nslots, grid, d_cta_first);
  MF_LAUNCH_CHECK();
  if (rel_count_bits == 32) {
    const size_t smem = count_stream2_smem_bytes<true>();
    set_smem(k_count_stream2<true, 32>, smem);
    k_count_stream2<true, 32><<<grid, kC2NT, smem, c.stream>>>(a, d_cta_first, 64 - key_bits, 32);
  } else {
    const size_t smem = count_stream2_smem_bytes<false>();
    set_smem(k_count_stream2<false, 0>, smem);
    k_count_stream2<false, 0><<<grid, kC2NT, smem, c.stream>>>(a, d_cta_first, 64 - key_bits, 0);
  }
  MF_LAUNCH_CHECK();
  c.launches += 2;
}

template <int W>
static void launch_count_stream_w(Ctx &c, const LocalArgs &a, int nslots, int grid, int32_t *d_cta_first) {
  if constexpr (W >= 3) {
    k_split_ranges<<<div_ceil(grid + 1, 128), 128, 0, c.stream>>>(a.bkt_start, a.bkt_size, nslots, grid, d_cta_first);
    MF_LAUNCH_CHECK();
    // W >= 7 (k >= 96): 2 ring stages of 512 keys (W = 9, 10: 384) instead of 3 of 256 -- every thread has a key in every chunk, still
    // two CTAs per SM (k=119: 50.7 -> 38.3 ms, k=99: 63.5 -> 51.5 ms, k=141: 20.4 -> 18.1 ms; MFSDBG_CW_CHUNK=0 for the old ring)
    {
      // the two-stage ring at the other widths (same or less shared memory than their three-stage rings): W = 3 and 5 gain 2-3 ms
      // (k = 39, 47, 69, 79), W = 4 nothing, W = 6 loses 6 ms with its 768-key chunks (profiles/r2ag_wide_ab_matrix.json)
      constexpr int CH2 = W == 3 ? 1536 : (W <= 5 ? 1024 : (W == 6 ? 768 : (W <= 8 ? 512 : 384)));
      const int want = env_int("MFSDBG_CW_CHUNK", -1);
      if (want == 1 || (want < 0 && (W >= 7 || W == 3 || W == 5))) {
        const size_t smem = count_stream_w_smem_bytes<W, CH2, 2>();
        set_smem(k_count_stream_w<W, CH2, 2>, smem);
        k_count_stream_w<W, CH2, 2><<<grid, kCwNT, smem, c.stream>>>(a, d_cta_first);
        MF_LAUNCH_CHECK();
        c.launches += 2;
        return;
      }
    }
    const size_t smem = count_stream_w_smem_bytes<W>();
    set_smem(k_count_stream_w<W>, smem);
    k_count_stream_w<W><<<grid, kCwNT, smem, c.stream>>>(a, d_cta_first);
    MF_LAUNCH_CHECK();
    c.launches += 2;
  }
}

// ------------------------------------------------------------------ count: finish from l1-partitioned keys
template <int W>
static void count_finish_impl(Ctx &c, uint32_t *cur, uint32_t *other, int64_t n, const HostChunks &l1, int k, int l1_bits,
                              int min_count, bool append, EdgesView *out, unsigned long long *d_counting) {
  const int key_bits = 2 * (k + 1), We = words_edge(k);
  Plan p = make_plan(W, key_bits, n, 2.0, true, l1_bits, l1.nseg);
  auto salloc = [&](size_t bytes) { return c.slab_alloc(bytes); };
  int bit_off = l1_bits;
  // keys of 33..64 bits take the streamed finish (TMA ring + shared hash table, buckets sized by a distinct-ratio probe)
  const bool stream = W >= 2 && env_int(W == 2 ? "MFSDBG_COUNT_STREAM" : "MFSDBG_COUNT_STREAM_W", 1) != 0 && !(d_counting && min_count > 64);
  DevBuckets b;
  double rho = 0.0;   // distinct / occurrences, when the streamed path measured it
  int rel_count_bits = 0;
  if constexpr (W >= 2) {
    if (stream) b = stream_partition<W>(c, &cur, &other, l1, &bit_off, key_bits, min_count, &rho, &rel_count_bits, salloc);
  }
  if (!stream) b = partition_chain<W>(c, &cur, &other, l1, &bit_off, p.rest_bits, salloc, "count_l2");
  const int stream_grid = stream ? (int)std::max<int64_t>(1, std::min<int64_t>(2 * c.sm_count, n / 16384)) : 0;
  int32_t *d_cta_first = stream ? c.alloc<int32_t>(stream_grid + 2) : nullptr;
  int64_t *d_desc_off = c.alloc<int64_t>(b.nslots), *d_desc_cnt = c.alloc<int64_t>(b.nslots);
  int64_t *d_out_off = c.alloc<int64_t>(b.nslots + 1);
  int32_t *d_bail = c.alloc<int32_t>(b.nslots);
  int *d_flags = c.alloc<int>(4);
  unsigned long long *d_cursor = c.alloc<unsigned long long>(2);
  // arena: what is left of the slab (edges are a small fraction of the keys unless -m 1)
  size_t arena_cap;
  {
    const size_t want_max = (size_t)(min_count > 1 ? n / min_count + 1 : n);
    size_t guess = std::min(want_max, (size_t)std::max<int64_t>(n / 6 + 1, 1 << 16));
    if (rho > 0.0 && min_count <= 1) guess = std::max(guess, std::min(want_max, (size_t)(1.15 * rho * (double)n) + 4096));
    guess += (size_t)stream_grid * kCsArenaBlock;
    const size_t room = c.slab_bytes - ((c.slab_off + 255) & ~(size_t)255);
    arena_cap = std::min(guess, room / ((size_t)We * 4));
  }
  uint32_t *d_arena = c.alloc<uint32_t>(arena_cap * We);
  DevBuf big_arena;   // only if the slab share overflowed
  bool fallback_sorted = false;
  for (int attempt = 0;; ++attempt) {
    MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * 4, c.stream));
    MF_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long) * 2, c.stream));
    MF_CUDA(cudaMemsetAsync(d_desc_cnt, 0, sizeof(int64_t) * b.nslots, c.stream));
    LocalArgs a{};
    a.in = cur;
    a.bkt_start = b.start;
    a.bkt_size = b.size;
    a.bit_off = bit_off;
    a.sort_bits = key_bits;
    a.cap = p.cap;
    a.k = k;
    a.min_count = min_count;
    a.words_edge = We;
    a.arena = d_arena;
    a.arena_cursor = d_cursor;
    a.arena_cap = arena_cap;
    a.desc_off = d_desc_off;
    a.desc_cnt = d_desc_cnt;
    a.counting = attempt == 0 ? d_counting : nullptr;   // a retry must not count the histogram twice
    a.bail_list = d_bail;
    a.bail_count = d_flags;
    a.overflow_flag = d_flags + 1;
    int flags[3];
    if constexpr (W <= 2) {
      // keys of <= 64 bits: stream every bucket through the key-resident hash table (any bucket size)
      {
        Stage st(c, "local_count");
        if (stream) launch_count_stream(c, a, b.nslots, n, stream_grid, d_cta_first, key_bits, rel_count_bits);
        else launch_count_fast<W>(c, a, b.nslots);
      }
      c.d2h(flags, d_flags, sizeof(int) * 3);
      if (flags[0] > 0) {
        // buckets with too many distinct keys for the table: the same kernel, one CTA per bucket, in passes over 8 sub-ranges
        if (stream) {
          Stage st(c, "local_count_multipass");
          std::vector<int32_t> slots;
          std::vector<Range> rs = fetch_bails(c, b, d_bail, flags[0], &slots);
          std::vector<WorkItem> wi(rs.size());
          for (size_t i = 0; i < rs.size(); ++i) wi[i] = WorkItem{rs[i].start, (int32_t)std::min<int64_t>(rs[i].size, INT32_MAX), slots[i]};
          c.ov[7].reserve(sizeof(WorkItem) * wi.size());
          c.h2d(c.ov[7].p, wi.data(), sizeof(WorkItem) * wi.size());
          MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), c.stream));
          LocalArgs ma = a;
          ma.work = c.ov[7].as<WorkItem>();
          const size_t smem = count_multipass_smem_bytes();
          set_smem(k_count_multipass, smem);
          k_count_multipass<<<(unsigned)wi.size(), kCsNT, smem, c.stream>>>(ma, 3);
          MF_LAUNCH_CHECK();
          c.launches++;
          if (getenv("MFSDBG_TRACE")) fprintf(stderr, "[mfsdbg] count: %d of %d buckets take the multi-pass kernel\n", (int)wi.size(), b.nslots);
          c.d2h(flags, d_flags, sizeof(int) * 3);
        }
      }
      if (flags[0] > 0) {
        // buckets that still do not fit: general path on exactly those slots
        Stage st(c, "local_count_general");
        if (getenv("MFSDBG_TRACE")) fprintf(stderr, "[mfsdbg] count: %d of %d buckets take the general kernel\n", flags[0], b.nslots);
        std::vector<int32_t> slots(flags[0]);
        c.d2h(slots.data(), d_bail, sizeof(int32_t) * flags[0]);
        std::sort(slots.begin(), slots.end());
        std::vector<WorkItem> wi;
        for (int sl : slots) wi.push_back(WorkItem{0, 0, sl});
        std::vector<Range> rs = fetch_bails(c, b, d_bail, flags[0], &slots);
        for (size_t i = 0; i < rs.size(); ++i) { wi[i].start = rs[i].start; wi[i].n = (int32_t)std::min<int64_t>(rs[i].size, INT32_MAX); }
        c.ov[7].reserve(sizeof(WorkItem) * wi.size());
        c.h2d(c.ov[7].p, wi.data(), sizeof(WorkItem) * wi.size());
        MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), c.stream));
        LocalArgs ga = a;
        ga.work = c.ov[7].as<WorkItem>();
        launch_local<W, kCountEmit>(c, ga, (int)wi.size());
        c.d2h(flags, d_flags, sizeof(int) * 3);
      }
    } else if (stream) {
      // keys wider than 64 bits: streamed finish with fingerprint + representative slots; what bails takes the general kernel
      {
        Stage st(c, "local_count");
        launch_count_stream_w<W>(c, a, b.nslots, stream_grid, d_cta_first);
      }
      c.d2h(flags, d_flags, sizeof(int) * 3);
      if (flags[0] > 0) {
        Stage st(c, "local_count_general");
        if (getenv("MFSDBG_TRACE")) fprintf(stderr, "[mfsdbg] count: %d of %d buckets take the general kernel\n", flags[0], b.nslots);
        std::vector<int32_t> slots;
        std::vector<Range> rs = fetch_bails(c, b, d_bail, flags[0], &slots);
        std::vector<WorkItem> wi(rs.size());
        for (size_t i = 0; i < rs.size(); ++i) wi[i] = WorkItem{rs[i].start, (int32_t)std::min<int64_t>(rs[i].size, INT32_MAX), slots[i]};
        c.ov[7].reserve(sizeof(WorkItem) * wi.size());
        c.h2d(c.ov[7].p, wi.data(), sizeof(WorkItem) * wi.size());
        MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), c.stream));
        LocalArgs ga = a;
        ga.work = c.ov[7].as<WorkItem>();
        launch_local<W, kCountEmit>(c, ga, (int)wi.size());
        c.d2h(flags, d_flags, sizeof(int) * 3);
      }
    } else {
      Stage st(c, "local_count");
      launch_local<W, kCountEmit>(c, a, b.nslots);
      c.d2h(flags, d_flags, sizeof(int) * 3);
    }
    if (flags[0] > 0) {
      // Buckets larger than shared memory.  With deep coverage (a mitogenome at 10^4 x) they are a few keys repeated
      // thousands of times: combine chunk by chunk into (key, count) pairs, then merge the pairs of each bucket.
      Stage st(c, "oversized");
      std::vector<int32_t> slots;
      std::vector<Range> rs = fetch_bails(c, b, d_bail, flags[0], &slots);
      std::vector<WorkItem> chunks, merges;
      // one CTA per SM for these few launches: larger chunks leave fewer (key, count) pairs for the merge, which in turn
      // holds more of them (W = 8: 5100-key chunks and 3600 pairs instead of 1600 and 520)
      const int big_cap = local_cap(W, true, false, kLocalBigSmem);
      for (size_t i = 0; i < rs.size(); ++i) {
        WorkItem m{(int64_t)chunks.size(), 0, slots[i]};
        for (int64_t off = 0; off < rs[i].size; off += big_cap) {
          chunks.push_back(WorkItem{rs[i].start + off, (int32_t)std::min<int64_t>(big_cap, rs[i].size - off), slots[i]});
          m.n++;
        }
        merges.push_back(m);
      }
      int64_t total = 0;
      for (auto &r : rs) total += r.size;
      DevBuf &d_chunks = c.ov[0], &d_merges = c.ov[1], &d_poff = c.ov[2], &d_pcnt = c.ov[3], &d_pairs = c.ov[4], &d_bail2 = c.ov[5],
             &d_pcur = c.ov[6];
      d_chunks.reserve(sizeof(WorkItem) * chunks.size());
      d_merges.reserve(sizeof(WorkItem) * merges.size());
      d_poff.reserve(sizeof(int64_t) * chunks.size());
      d_pcnt.reserve(sizeof(int32_t) * chunks.size());
      d_bail2.reserve(sizeof(int32_t) * merges.size());
      d_pcur.reserve(64);
      c.h2d(d_chunks.p, chunks.data(), sizeof(WorkItem) * chunks.size());
      c.h2d(d_merges.p, merges.data(), sizeof(WorkItem) * merges.size());
      size_t pair_cap = (size_t)std::min<int64_t>(total, total / 8 + (1 << 20));
      std::vector<int32_t> hard;   // buckets whose DISTINCT keys do not fit either
      for (int ptry = 0;; ++ptry) {
        d_pairs.reserve(pair_cap * (W + 1) * 4);
        MF_CUDA(cudaMemsetAsync(d_pcur.p, 0, 64, c.stream));
        MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * 4, c.stream));
        LocalArgs pa = a;
        pa.cap = big_cap;
        pa.work = d_chunks.as<WorkItem>();
        pa.counting = nullptr;
        pa.pair_arena = d_pairs.as<uint32_t>();
        pa.pair_cursor = d_pcur.as<unsigned long long>();
        pa.pair_cap = pair_cap;
        pa.piece_off = d_poff.as<int64_t>();
        pa.piece_cnt = d_pcnt.as<int32_t>();
        {
          Stage st2(c, "oversized_pairs");
          launch_local<W, kCountPairs>(c, pa, (int)chunks.size());
        }
        int pf[3];
        c.d2h(pf, d_flags, sizeof(int) * 3);
        if (pf[2]) {   // pair arena too small: every key distinct is the bound
          if (ptry) throw std::runtime_error("pair arena overflow persisted");
          pair_cap = (size_t)total;
          continue;
        }
        LocalArgs ma = pa;
        ma.work = d_merges.as<WorkItem>();
        ma.cap = local_cap(W + 1, true, true, kLocalBigSmem);
        ma.counting = a.counting;
        ma.bail_list = d_bail2.as<int32_t>();
        {
          Stage st2(c, "oversized_merge");
          launch_local<W, kCountMerge>(c, ma, (int)merges.size());
        }
        c.d2h(pf, d_flags, sizeof(int) * 3);
        flags[1] |= pf[1];
        if (pf[0] > 0) {
          hard.resize(pf[0]);
          c.d2h(hard.data(), d_bail2.p, sizeof(int32_t) * pf[0]);
        }
        break;
      }
      if (!hard.empty()) {
        std::sort(hard.begin(), hard.end());
        std::vector<Range> hr;
        std::vector<WorkItem> hw;
        for (size_t i = 0; i < rs.size(); ++i)
          if (std::binary_search(hard.begin(), hard.end(), slots[i])) {
            hr.push_back(rs[i]);
            hw.push_back(WorkItem{rs[i].start, 0, slots[i]});
          }
        if (getenv("MFSDBG_TRACE")) {
          int64_t tot = 0, mx = 0;
          for (auto &r : hr) { tot += r.size; mx = std::max(mx, r.size); }
          fprintf(stderr, "[mfsdbg] count: %d of %d oversized buckets take the sorted-run path (%lld keys, largest %lld)\n", (int)hr.size(),
                  (int)rs.size(), (long long)tot, (long long)mx);
        }
        if (!fallback_sorted) {
          Stage st2(c, "oversized_sort");
          sort_ranges<W>(c, cur, other, hr, bit_off, key_bits);
        }
        fallback_sorted = true;
        DevBuf &d_hw = c.ov[7];
        d_hw.reserve(sizeof(WorkItem) * hw.size());
        c.h2d(d_hw.p, hw.data(), sizeof(WorkItem) * hw.size());
        LocalArgs sa = a;
        sa.work = d_hw.as<WorkItem>();
        {
          Stage st2(c, "oversized_runs");
          k_sorted_runs<W, 256><<<(unsigned)hw.size(), 256, 0, c.stream>>>(sa, (int)hw.size());
          MF_LAUNCH_CHECK();
          c.launches++;
        }
        int sf[3];
        c.d2h(sf, d_flags, sizeof(int) * 3);
        flags[1] |= sf[1];
      }
    }
    if (!flags[1]) break;
    // arena overflow: the true demand is in the cursor; take it from a dedicated allocation and redo the finish
    unsigned long long need = 0;
    c.d2h(&need, d_cursor, sizeof need);
    if (attempt >= 2) throw std::runtime_error("edge arena overflow persisted");
    big_arena.reserve((size_t)need * We * 4);
    d_arena = big_arena.as<uint32_t>();
    arena_cap = need;
  }
  // gather
  Stage st(c, "gather");
  scan_slots(c, d_desc_cnt, b.nslots, d_out_off);
  int64_t E = 0;
  MF_CUDA(cudaMemcpyAsync(&E, d_out_off + b.nslots, sizeof E, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  const int64_t prev = append ? out->n_edges : 0;
  if (append && prev > 0) {
    // out-of-core rounds append: grow geometrically (a fresh exact-size allocation + copy per round cost 577 ms of the 1.54 s
    // of the 30 Gbp run: cudaMalloc / cudaFree of multi-GB blocks stall, and the copies are quadratic in the rounds)
    const size_t need = (size_t)(prev + E) * We * 4 + 256;
    if (need > c.edges.cap) {
      DevBuf grown;
      grown.reserve(std::max(need, c.edges.cap * 2));
      MF_CUDA(cudaMemcpyAsync(grown.p, c.edges.p, (size_t)prev * We * 4, cudaMemcpyDeviceToDevice, c.stream));
      MF_CUDA(cudaStreamSynchronize(c.stream));
      c.edges.release();
      c.edges = grown;
    }
  } else {
    c.edges.reserve((size_t)std::max<int64_t>(E, 1) * We * 4 + 256);
  }
  uint32_t *d_edges = c.edges.as<uint32_t>() + prev * We;
  if (E > 0) {
    k_gather_edges<<<div_ceil(b.nslots, 8), 256, 0, c.stream>>>(d_arena, d_desc_off, d_desc_cnt, d_out_off, b.nslots, We, d_edges);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  MF_CUDA(cudaStreamSynchronize(c.stream));
  big_arena.release();
  out->edges = c.edges.as<uint32_t>();
  out->n_edges = prev + E;
  out->k = k;
  out->words = We;
}

static void edge_bucket_counts(Ctx &c, const EdgesView &e) {
  c.small[0].reserve(sizeof(unsigned long long) * kNumBuckets);
  unsigned long long *d = c.small[0].as<unsigned long long>();
  MF_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * kNumBuckets, c.stream));
  if (e.n_edges > 0) {
    k_edge_buckets<<<kNumBuckets / 256, 256, 0, c.stream>>>(e.edges, e.n_edges, e.words, d);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  MF_CUDA(cudaMemcpyAsync(c.edge_bucket_counts.data(), d, sizeof(int64_t) * kNumBuckets, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
}

static const uint32_t *build_start_bits(Ctx &c, const ReadsView &r) {
  size_t words = (size_t)((r.n_bases + 31) >> 5) + 64;
  c.sbits.reserve(words * 4);
  MF_CUDA(cudaMemsetAsync(c.sbits.p, 0, words * 4, c.stream));
  if (r.n_reads > 0) {
    k_start_bits<<<(unsigned)div_ceil64(r.n_reads, 256), 256, 0, c.stream>>>(r.starts, r.n_reads, c.sbits.as<uint32_t>());
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  return c.sbits.as<uint32_t>();
}

const uint32_t *dev_start_bits(Ctx &c, const ReadsView &r) { return build_start_bits(c, r); }

// level-1 digit width of count
template <int W>
static Plan count_plan(Ctx &c, int k, int min_count, int64_t n_est, bool compact = false) {
  const int key_bits = 2 * (k + 1);
  Plan p = make_plan(W, key_bits, n_est, 2.0, true);
  if (W >= 2 && env_int("MFSDBG_L1_BITS", -1) < 0 && env_int(W == 2 ? "MFSDBG_COUNT_STREAM" : "MFSDBG_COUNT_STREAM_W", 1) != 0) {
    // streamed finish: level 2 is a range partition into <= 1024 bins of ~16 K keys (45 % table load at the usual distinct
    // ratio), so level 1 only has to bring the densest segments (2x the average) under ~17-24 M keys -- and the fewer bins the
    // reads-fed scatter has, the faster it runs (512 bins: 14.9 ms, 1024: 19.4 ms, 2048: 48 ms on the 5 Gbp sample)
    // keys of 4+ words are staged in small tiles (ReadsTileCfg: 256 / 128 threads, 4096 / 2048 positions, and only (L-k)/L
    // of the positions start a key): beyond 256 bins a tile holds less than a key or two per bin and the reads-fed scatter
    // falls off a cliff (k=79: 51 ms at 8 bits, 96 ms at 9); the extra bits go to the prefix level that follows anyway
    // (the compacted scatter stages the keys of 8192 positions whatever their width and takes up to 1024 bins:
    // MFSDBG_L1_CAP_WC / MFSDBG_L1_SEGK_WC = its bit cap and the keys, in thousands, a level-1 segment should hold)
    const int l1_cap = W >= 4 ? (compact ? std::min(10, env_int("MFSDBG_L1_CAP_WC", 10)) : 8) : 10;
    const double seg_keys = W == 2 ? 1.2e7 : (compact ? 1e3 * std::max(1, env_int("MFSDBG_L1_SEGK_WC", 1000)) : 4.0e6);
    p.l1_bits = std::max(1, std::min({l1_cap, key_bits, ceil_log2((double)n_est / seg_keys)}));
    // out-of-core rounds are cut at level-1 bin boundaries: a bin (twice the average at small prefixes) must stay well
    // inside what one round may hold
    const size_t bud0 = (size_t)((double)c.budget() * 0.9);
    const double pk0 = 2.0 * W * 4 + (min_count > 1 ? (double)words_edge(k) * 4 / std::min(min_count, 6) : (double)words_edge(k) * 4);
    while (p.l1_bits < std::min(kMaxDigitBits, key_bits)) {
      const size_t tb0 = (size_t)(64 << 20) + (size_t)(((size_t)1 << p.l1_bits) << kMaxDigitBits) * 96;
      const double mk0 = bud0 > tb0 ? (double)(bud0 - tb0) / pk0 : 0.0;
      if (mk0 >= (double)n_est || 3.0 * (double)n_est / (double)(1 << p.l1_bits) <= mk0) break;
      ++p.l1_bits;
    }
  }
  return p;
}

template <int W>
static void dev_count_impl(Ctx &c, const ReadsView &r, int k, int min_count, EdgesView *out, int64_t *counting_host) {
  const int key_bits = 2 * (k + 1), We = words_edge(k);
  const uint32_t *sbits;
  {
    Stage st(c, "start_bits");
    sbits = build_start_bits(c, r);
  }
  // the plan needs the key count: it is at most one key per base
  const int64_t n_est = std::max<int64_t>(r.n_bases - r.n_reads * (int64_t)k, 1);   // one key per base minus k per read
  Plan p = count_plan<W>(c, k, min_count, n_est, reads_compact_wanted(W, r.n_bases, r.n_reads, k));
  const int nb1 = 1 << p.l1_bits;
  c.small[1].reserve(sizeof(unsigned long long) * (kMaxBins + kNumBuckets));
  unsigned long long *d_small = c.small[1].as<unsigned long long>();   // hist | counting
  unsigned long long *d_hist = d_small, *d_counting = d_small + nb1;
  MF_CUDA(cudaMemsetAsync(d_small, 0, sizeof(unsigned long long) * (nb1 + kNumBuckets), c.stream));
  // ---- sampled histogram: every 64th tile is histogrammed, every bin gets its scaled share plus 10 % and the scatter
  // keeps exact cursors; a bin that outgrows its region (cursor > limit) sends the call down the exact path below.  Saves
  // the full key-extraction pass over the reads (4.7 ms of the 5 Gbp step); a strided sample sees the whole file, so reads
  // sorted by position do not fool it.
  {
    using RC = ReadsTileCfg<W>;
    const int64_t ntiles = div_ceil64(r.n_bases, RC::T), stride = std::max(1, env_int("MFSDBG_SAMPLED_STRIDE", 64)), nsamp = div_ceil64(ntiles, stride);
    const size_t tb0 = (size_t)(64 << 20) + (size_t)((size_t)nb1 << kMaxDigitBits) * 96;
    const size_t bud0 = (size_t)((double)c.budget() * 0.9);
    const double pk0 = 2.0 * W * 4 + (min_count > 1 ? (double)We * 4 / std::min(min_count, 6) : (double)We * 4);
    const double cap_total = 1.12 * (double)n_est + 131072.0 * nb1;
    if (W >= 2 && env_int("MFSDBG_SAMPLED_HIST", 1) != 0 && ntiles >= (int64_t)env_int("MFSDBG_SAMPLED_MIN_TILES", 4096) && cap_total * pk0 + (double)tb0 < (double)bud0) {
      {
        Stage st(c, "reads_hist");
        launch_reads_hist<W>(c, r, sbits, k, LevelArgs{0, p.l1_bits, 0u, (uint32_t)nb1, nullptr}, d_hist, 0, nsamp, stride);
      }
      std::vector<unsigned long long> hs(nb1);
      c.d2h(hs.data(), d_hist, sizeof(unsigned long long) * nb1);
      const double scale = (double)ntiles / (double)nsamp;
      std::vector<unsigned long long> start(nb1), limit(nb1);
      unsigned long long acc = 0;
      for (int b = 0; b < nb1; ++b) {
        start[b] = acc;
        acc += (unsigned long long)((double)hs[b] * scale * (1.0 + env_int("MFSDBG_SAMPLED_SLACK_PCT", 10) / 100.0)) + 65536ull;
        acc = (acc + 1ull) & ~1ull;   // regions start on 16-byte boundaries
        limit[b] = acc;
      }
      if ((double)acc <= cap_total) {
        c.slab_reserve((size_t)((double)acc * pk0) + tb0 + (1 << 20));
        c.slab_reset();
        uint32_t *bufA = c.alloc<uint32_t>((size_t)acc * W + 64), *bufB = c.alloc<uint32_t>((size_t)acc * W + 64);
        unsigned long long *d_cursor = c.alloc<unsigned long long>(nb1), *d_limit = c.alloc<unsigned long long>(nb1);
        c.h2d(d_cursor, start.data(), sizeof(unsigned long long) * nb1);
        c.h2d(d_limit, limit.data(), sizeof(unsigned long long) * nb1);
        {
          Stage st(c, "reads_scatter");
          LevelArgs la{0, p.l1_bits, 0u, (uint32_t)nb1, nullptr};
          la.limit = d_limit;
          launch_reads_scatter<W>(c, r, sbits, k, la, d_cursor, bufA);
        }
        std::vector<unsigned long long> cur(nb1);
        c.d2h(cur.data(), d_cursor, sizeof(unsigned long long) * nb1);
        bool fits = true;
        for (int b = 0; b < nb1; ++b) fits = fits && cur[b] <= limit[b];
        if (fits) {
          HostChunks l1;
          l1.nseg = nb1;
          int64_t n_keys = 0;
          for (int b = 0; b < nb1; ++b) {
            const int64_t sz = (int64_t)(cur[b] - start[b]);
            l1.start.push_back((int64_t)start[b]);
            l1.size.push_back(sz);
            l1.seg.push_back(b);
            l1.seg_out_start.push_back(n_keys);
            n_keys += sz;
          }
          out->n_keys = n_keys;
          out->n_edges = 0;
          out->k = k;
          out->words = We;
          if (n_keys > 0) count_finish_impl<W>(c, bufA, bufB, n_keys, l1, k, p.l1_bits, min_count, false, out, counting_host ? d_counting : nullptr);
          if (out->n_edges == 0) {
            c.edges.reserve(256);
            out->edges = c.edges.as<uint32_t>();
          }
          if (counting_host) c.d2h(counting_host, d_counting, sizeof(int64_t) * kNumBuckets);
          Stage st(c, "edge_buckets");
          edge_bucket_counts(c, *out);
          return;
        }
        if (getenv("MFSDBG_TRACE")) fprintf(stderr, "[mfsdbg] sampled histogram: a bin outgrew its region, exact pass instead\n");
        { Stage mark(c, "sampled_overflow"); }   // shows up in the call's profile (tests look for it)
      }
      MF_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nb1, c.stream));
    }
  }
  {
    Stage st(c, "reads_hist");
    launch_reads_hist<W>(c, r, sbits, k, LevelArgs{0, p.l1_bits, 0u, (uint32_t)nb1, nullptr}, d_hist);
  }
  std::vector<unsigned long long> hist(nb1);
  MF_CUDA(cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned long long) * nb1, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  int64_t n_keys = 0;
  for (auto h : hist) n_keys += (int64_t)h;
  out->n_keys = n_keys;
  out->n_edges = 0;
  out->k = k;
  out->words = We;
  // rounds over contiguous l1-bin ranges so that two key buffers + tables fit the budget
  const size_t table_bytes = (size_t)(64 << 20) + (size_t)((size_t)nb1 << kMaxDigitBits) * 96;
  const size_t budget = (size_t)((double)c.budget() * 0.9);   // a tenth is left for the resulting edges / sdbg
  const double per_key = 2.0 * W * 4 + (min_count > 1 ? (double)We * 4 / std::min(min_count, 6) : (double)We * 4);
  int64_t max_keys = budget > table_bytes ? (int64_t)((double)(budget - table_bytes) / per_key) : 0;
  if (max_keys < 1024) throw std::runtime_error("device memory budget too small for count");
  std::vector<std::pair<int, int>> rounds;
  for (int lo = 0; lo < nb1;) {
    int hi = lo;
    int64_t acc = 0;
    while (hi < nb1 && (acc + (int64_t)hist[hi] <= max_keys || hi == lo)) acc += (int64_t)hist[hi++];
    if (acc > max_keys) throw std::runtime_error("a single prefix bin exceeds the device memory budget");
    rounds.emplace_back(lo, hi);
    lo = hi;
  }
  int64_t round_max = 0;
  for (auto &rd : rounds) {
    int64_t acc = 0;
    for (int b = rd.first; b < rd.second; ++b) acc += (int64_t)hist[b];
    round_max = std::max(round_max, acc);
  }
  c.slab_reserve((size_t)((double)round_max * per_key) + table_bytes + (1 << 20));
  for (size_t ri = 0; ri < rounds.size(); ++ri) {
    const int lo = rounds[ri].first, hi = rounds[ri].second;
    c.slab_reset();
    HostChunks l1;
    l1.nseg = hi - lo;
    std::vector<unsigned long long> cursor(nb1, 0);
    int64_t acc = 0;
    for (int b = lo; b < hi; ++b) {
      cursor[b] = (unsigned long long)acc;
      l1.start.push_back(acc);
      l1.size.push_back((int64_t)hist[b]);
      l1.seg.push_back(b - lo);
      l1.seg_out_start.push_back(acc);
      acc += (int64_t)hist[b];
    }
    if (acc == 0) continue;
    uint32_t *bufA = c.alloc<uint32_t>((size_t)acc * W + 64), *bufB = c.alloc<uint32_t>((size_t)acc * W + 64);   // bulk copies round up to 16 bytes
    unsigned long long *d_cursor = c.alloc<unsigned long long>(nb1);
    c.h2d(d_cursor, cursor.data(), sizeof(unsigned long long) * nb1);
    {
      Stage st(c, "reads_scatter");
      launch_reads_scatter<W>(c, r, sbits, k, LevelArgs{0, p.l1_bits, (uint32_t)lo, (uint32_t)hi, nullptr}, d_cursor, bufA);
    }
    count_finish_impl<W>(c, bufA, bufB, acc, l1, k, p.l1_bits, min_count, ri > 0, out, counting_host ? d_counting : nullptr);
  }
  if (out->n_edges == 0) {
    c.edges.reserve(256);
    out->edges = c.edges.as<uint32_t>();
  }
  if (counting_host)
    c.d2h(counting_host, d_counting, sizeof(int64_t) * kNumBuckets);
  Stage st(c, "edge_buckets");
  edge_bucket_counts(c, *out);
}

#define MF_DISPATCH_CASE_COUNT(Wn) \
  case Wn: dev_count_impl<Wn>(c, r, k, min_count, out, counting_host); break;
void dev_count(Ctx &c, const ReadsView &r, int k, int min_count, EdgesView *out, int64_t *counting_host) {
  if (k < 9 || k > 150) throw std::invalid_argument("k must be in [9, 150]");
  if (min_count < 1) throw std::invalid_argument("min_count must be >= 1");
  MF_DISPATCH_W(words_key(k), COUNT)
}


// ------------------------------------------------------------------ count with host-resident reads (pipelined H2D)
template <int W>
static bool dev_count_host_impl(Ctx &c, const uint32_t *packed_host, const int64_t *starts_host, int64_t n_reads, int64_t n_bases, int k,
                                int min_count, EdgesView *out) {
  using C = ReadsTileCfg<W>;
  const int We = words_edge(k);
  const int64_t total_words = (n_bases + 15) >> 4;
  const int64_t ntiles = div_ceil64(n_bases, C::T);
  const int n_chunks = (int)std::min<int64_t>(env_int("MFSDBG_H2D_CHUNKS", 8), ntiles / 8);
  if (n_chunks < 2) return false;
  const int64_t n_est = std::max<int64_t>(n_bases - n_reads * (int64_t)k, 1);
  Plan p = count_plan<W>(c, k, min_count, n_est, reads_compact_wanted(W, n_bases, n_reads, k));
  const int nb1 = 1 << p.l1_bits;
  const size_t table_bytes = (size_t)(64 << 20) + (size_t)((size_t)nb1 << kMaxDigitBits) * 96 + sizeof(int64_t) * 8 * (size_t)nb1 * n_chunks;
  const size_t budget = (size_t)((double)c.budget() * 0.9);
  const double per_key = 2.0 * W * 4 + (min_count > 1 ? (double)We * 4 / std::min(min_count, 6) : (double)We * 4);
  const int64_t n_bound = n_bases;   // a key per base: reads shorter than k make n_est an estimate, not a bound
  if ((double)n_bound * per_key + (double)table_bytes > (double)budget) return false;   // needs out-of-core rounds
  const size_t wbytes = (size_t)total_words * 4, sbytes = sizeof(int64_t) * (size_t)(n_reads + 1);
  c.in_words.reserve(wbytes + 1024);
  c.in_starts.reserve(sbytes);
  const size_t sb_words = (size_t)((n_bases + 31) >> 5) + 64;
  c.sbits.reserve(sb_words * 4);
  if (!c.copy_stream) MF_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  while ((int)c.copy_events.size() < n_chunks + 1) {
    cudaEvent_t e;
    MF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c.copy_events.push_back(e);
  }
  c.slab_reserve((size_t)((double)n_bound * per_key) + table_bytes + (1 << 20));
  c.slab_reset();
  uint32_t *bufA = c.alloc<uint32_t>((size_t)n_bound * W + 64), *bufB = c.alloc<uint32_t>((size_t)n_bound * W + 64);
  unsigned long long *d_hist = c.alloc<unsigned long long>((size_t)nb1 * n_chunks);
  unsigned long long *d_cursor = c.alloc<unsigned long long>((size_t)nb1 * n_chunks);
  // everything the copy stream overwrites must be idle, and the start bits clean, before the first chunk lands
  MF_CUDA(cudaMemsetAsync(c.sbits.p, 0, sb_words * 4, c.stream));
  MF_CUDA(cudaMemsetAsync((char *)c.in_words.p + wbytes, 0, 1024, c.stream));
  MF_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nb1 * n_chunks, c.stream));
  MF_CUDA(cudaEventRecord(c.copy_events[n_chunks], c.stream));
  MF_CUDA(cudaStreamWaitEvent(c.copy_stream, c.copy_events[n_chunks], 0));
  const int64_t tiles_per = div_ceil64(ntiles, n_chunks);
  struct Chunk { int64_t t0, t1, r_lo, r_hi; };
  std::vector<Chunk> chunks;
  const int64_t margin = k + 256;   // a key near the chunk end looks this far ahead for read starts and bases
  for (int ci = 0; ci < n_chunks; ++ci) {
    Chunk ch;
    ch.t0 = std::min<int64_t>(ntiles, ci * tiles_per);
    ch.t1 = std::min<int64_t>(ntiles, (ci + 1) * tiles_per);
    if (ch.t1 <= ch.t0) break;
    const int64_t b0 = ch.t0 * C::T, b1 = std::min<int64_t>(n_bases, ch.t1 * C::T);
    const int64_t w0 = b0 >> 4, w1 = std::min<int64_t>(total_words, ((b1 + margin) >> 4) + 2);
    // words already sent with the previous chunk's margin are sent again (a few hundred bytes)
    MF_CUDA(cudaMemcpyAsync(c.in_words.as<uint32_t>() + w0, packed_host + w0, (size_t)(w1 - w0) * 4, cudaMemcpyHostToDevice, c.copy_stream));
    ch.r_lo = std::lower_bound(starts_host, starts_host + n_reads + 1, b0) - starts_host;
    ch.r_hi = std::upper_bound(starts_host, starts_host + n_reads + 1, b1 + margin) - starts_host;
    if (ch.r_hi > ch.r_lo)
      MF_CUDA(cudaMemcpyAsync(c.in_starts.as<int64_t>() + ch.r_lo, starts_host + ch.r_lo, (size_t)(ch.r_hi - ch.r_lo) * 8, cudaMemcpyHostToDevice,
                              c.copy_stream));
    MF_CUDA(cudaEventRecord(c.copy_events[chunks.size()], c.copy_stream));
    chunks.push_back(ch);
  }
  ReadsView r{c.in_words.as<uint32_t>(), c.in_starts.as<int64_t>(), n_reads, n_bases};
  const uint32_t *sbits = c.sbits.as<uint32_t>();
  HostChunks l1;
  l1.nseg = nb1;
  std::vector<int64_t> seg_total(nb1, 0);
  std::vector<unsigned long long> hist(nb1), cursor(nb1);
  int64_t acc = 0;
  for (size_t ci = 0; ci < chunks.size(); ++ci) {
    const Chunk &ch = chunks[ci];
    MF_CUDA(cudaStreamWaitEvent(c.stream, c.copy_events[ci], 0));
    const int64_t nr = std::min<int64_t>(ch.r_hi, n_reads) - ch.r_lo;   // starts[n_reads] is the end sentinel, not a read
    if (nr > 0) {
      k_start_bits<<<(unsigned)div_ceil64(nr, 256), 256, 0, c.stream>>>(r.starts + ch.r_lo, nr, c.sbits.as<uint32_t>());
      MF_LAUNCH_CHECK();
      c.launches++;
    }
    unsigned long long *dh = d_hist + (size_t)ci * nb1, *dc = d_cursor + (size_t)ci * nb1;
    {
      Stage st(c, "reads_hist");
      launch_reads_hist<W>(c, r, sbits, k, LevelArgs{0, p.l1_bits, 0u, (uint32_t)nb1, nullptr}, dh, ch.t0, ch.t1 - ch.t0);
    }
    c.d2h(hist.data(), dh, sizeof(unsigned long long) * nb1);
    for (int b = 0; b < nb1; ++b) {
      cursor[b] = (unsigned long long)acc;
      if (hist[b]) {
        l1.start.push_back(acc);
        l1.size.push_back((int64_t)hist[b]);
        l1.seg.push_back(b);
      }
      seg_total[b] += (int64_t)hist[b];
      acc += (int64_t)hist[b];
    }
    if (acc > n_bound) throw std::runtime_error("key count exceeds its bound (internal error)");
    MF_CUDA(cudaMemcpyAsync(dc, cursor.data(), sizeof(unsigned long long) * nb1, cudaMemcpyHostToDevice, c.stream));
    MF_CUDA(cudaStreamSynchronize(c.stream));   // `cursor` is reused by the next chunk
    {
      Stage st(c, "reads_scatter");
      launch_reads_scatter<W>(c, r, sbits, k, LevelArgs{0, p.l1_bits, 0u, (uint32_t)nb1, nullptr}, dc, bufA, ch.t0, ch.t1 - ch.t0);
    }
  }
  {
    int64_t run = 0;
    for (int b = 0; b < nb1; ++b) { l1.seg_out_start.push_back(run); run += seg_total[b]; }
  }
  out->n_keys = acc;
  out->n_edges = 0;
  out->k = k;
  out->words = We;
  if (acc > 0) count_finish_impl<W>(c, bufA, bufB, acc, l1, k, p.l1_bits, min_count, false, out, nullptr);
  if (out->n_edges == 0) {
    c.edges.reserve(256);
    out->edges = c.edges.as<uint32_t>();
  }
  Stage st(c, "edge_buckets");
  edge_bucket_counts(c, *out);
  return true;
}
#define MF_DISPATCH_CASE_COUNTHOST(Wn) \
  case Wn: return dev_count_host_impl<Wn>(c, packed_host, starts_host, n_reads, n_bases, k, min_count, out);
bool dev_count_host(Ctx &c, const uint32_t *packed_host, const int64_t *starts_host, int64_t n_reads, int64_t n_bases, int k,
                    int min_count, EdgesView *out) {
  if (k < 9 || k > 150) throw std::invalid_argument("k must be in [9, 150]");
  if (min_count < 1) throw std::invalid_argument("min_count must be >= 1");
  MF_DISPATCH_W(words_key(k), COUNTHOST)
  return false;
}

// staged API for the multi-GPU driver
template <int W>
static void dev_count_hist_impl(Ctx &c, const ReadsView &r, int k, int l1_bits, unsigned long long *hist_dev) {
  const uint32_t *sbits = build_start_bits(c, r);
  const int nb1 = 1 << l1_bits;
  MF_CUDA(cudaMemsetAsync(hist_dev, 0, sizeof(unsigned long long) * nb1, c.stream));
  Stage st(c, "reads_hist");
  launch_reads_hist<W>(c, r, sbits, k, LevelArgs{0, l1_bits, 0u, (uint32_t)nb1, nullptr}, hist_dev);
}
#define MF_DISPATCH_CASE_CHIST(Wn) \
  case Wn: dev_count_hist_impl<Wn>(c, r, k, l1_bits, hist_dev); break;
void dev_count_hist(Ctx &c, const ReadsView &r, int k, int l1_bits, unsigned long long *hist_dev) {
  if (k < 9 || k > 150) throw std::invalid_argument("k must be in [9, 150]");
  if (l1_bits < 1 || l1_bits > kMaxDigitBits) throw std::invalid_argument("l1_bits must be in [1, 11]");
  MF_DISPATCH_W(words_key(k), CHIST)
}
__global__ void k_excl_scan_u64_small(const unsigned long long *in, int n, unsigned long long *out) {
  // n <= kMaxBins, one block
  __shared__ unsigned long long s[kMaxBins];
  for (int t = threadIdx.x; t < n; t += blockDim.x) s[t] = in[t];
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < n; ++i) { unsigned long long v = s[i]; s[i] = run; run += v; }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = s[t];
}
__global__ void k_sum_u64(const unsigned long long *in, int n, unsigned long long *out) {
  unsigned long long v = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += in[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}
template <int W>
static void dev_count_scatter_impl(Ctx &c, const ReadsView &r, int k, int l1_bits, const unsigned long long *hist_dev,
                                   uint32_t *keys_out, int64_t capacity, const unsigned long long *bin_base) {
  // the start bitmap is rebuilt for THESE reads: a cached one may belong to another reads object or an earlier call
  const uint32_t *sbits = build_start_bits(c, r);
  const int nb1 = 1 << l1_bits;
  c.slab_reserve(1 << 20);
  unsigned long long *d_cursor = c.alloc<unsigned long long>(nb1 + 1);
  if (bin_base) {
    MF_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long) * nb1, c.stream));   // offsets within each destination
  } else {
    // the histogram decides how much is written: it must fit what the caller allocated
    MF_CUDA(cudaMemsetAsync(d_cursor + nb1, 0, sizeof(unsigned long long), c.stream));
    k_sum_u64<<<1, 256, 0, c.stream>>>(hist_dev, nb1, d_cursor + nb1);
    k_excl_scan_u64_small<<<1, 1024, 0, c.stream>>>(hist_dev, nb1, d_cursor);
    MF_LAUNCH_CHECK();
    c.launches += 2;
    unsigned long long total = 0;
    c.d2h(&total, d_cursor + nb1, sizeof total);
    if (capacity >= 0 && total > (unsigned long long)capacity)
      throw std::invalid_argument("count_scatter: the histogram holds " + std::to_string(total) + " keys, keys_out has room for " +
                                  std::to_string(capacity));
  }
  Stage st(c, "reads_scatter");
  launch_reads_scatter<W>(c, r, sbits, k, LevelArgs{0, l1_bits, 0u, (uint32_t)nb1, bin_base}, d_cursor, keys_out);
}
#define MF_DISPATCH_CASE_CSCAT(Wn) \
  case Wn: dev_count_scatter_impl<Wn>(c, r, k, l1_bits, hist_dev, keys_out, capacity, bin_base); break;
void dev_count_scatter(Ctx &c, const ReadsView &r, int k, int l1_bits, const unsigned long long *hist_dev, uint32_t *keys_out,
                       int64_t capacity, const unsigned long long *bin_base) {
  if (k < 9 || k > 150) throw std::invalid_argument("k must be in [9, 150]");
  if (l1_bits < 1 || l1_bits > kMaxDigitBits) throw std::invalid_argument("l1_bits must be in [1, 11]");
  if (!bin_base && !hist_dev) throw std::invalid_argument("count_scatter needs a histogram or peer bin bases");
  MF_DISPATCH_W(words_key(k), CSCAT)
}
template <int W>
static void dev_count_finish_w(Ctx &c, uint32_t *keys, uint32_t *scratch, int64_t n_keys, const HostChunks &hc, int k, int l1_bits,
                               int min_count, EdgesView *out, int64_t *counting_host) {
  const int We = words_edge(k);
  c.small[2].reserve(sizeof(unsigned long long) * kNumBuckets);
  unsigned long long *d_counting = c.small[2].as<unsigned long long>();
  MF_CUDA(cudaMemsetAsync(d_counting, 0, sizeof(unsigned long long) * kNumBuckets, c.stream));
  const size_t table_bytes = (size_t)(64 << 20) + (size_t)((size_t)hc.nseg << kMaxDigitBits) * 96 + (size_t)(n_keys / 16);
  const size_t arena = (size_t)(min_count > 1 ? n_keys / std::min(min_count, 6) + 1 : n_keys) * We * 4;
  c.slab_reserve(table_bytes + arena + (1 << 20));
  out->n_edges = 0;
  out->n_keys = n_keys;
  out->k = k;
  out->words = We;
  if (n_keys > 0) count_finish_impl<W>(c, keys, scratch, n_keys, hc, k, l1_bits, min_count, false, out, counting_host ? d_counting : nullptr);
  if (out->n_edges == 0) {
    c.edges.reserve(256);
    out->edges = c.edges.as<uint32_t>();
  }
  if (counting_host) c.d2h(counting_host, d_counting, sizeof(int64_t) * kNumBuckets);
  edge_bucket_counts(c, *out);
}
#define MF_DISPATCH_CASE_CFIN(Wn) \
  case Wn: dev_count_finish_w<Wn>(c, keys, scratch, n_keys, hc, k, l1_bits, min_count, out, counting_host); break;
void dev_count_finish(Ctx &c, uint32_t *keys, uint32_t *scratch, int64_t n_keys, const int64_t *chunk_start,
                      const int64_t *chunk_size, const int32_t *chunk_seg, int n_chunks, int n_segs, int k, int l1_bits,
                      int min_count, EdgesView *out, int64_t *counting_host) {
  HostChunks hc;
  hc.nseg = n_segs;
  hc.start.assign(chunk_start, chunk_start + n_chunks);
  hc.size.assign(chunk_size, chunk_size + n_chunks);
  hc.seg.assign(chunk_seg, chunk_seg + n_chunks);
  std::vector<int64_t> tot(n_segs, 0);
  for (int i = 0; i < n_chunks; ++i) {
    if (chunk_seg[i] < 0 || chunk_seg[i] >= n_segs) throw std::invalid_argument("chunk_seg out of range");
    tot[chunk_seg[i]] += chunk_size[i];
  }
  int64_t acc = 0;
  for (int s = 0; s < n_segs; ++s) { hc.seg_out_start.push_back(acc); acc += tot[s]; }
  if (acc != n_keys) throw std::invalid_argument("chunk sizes do not add up to n_keys");
  MF_DISPATCH_W(words_key(k), CFIN)
}

// ------------------------------------------------------------------ seq2sdbg
// ---- the item filter (kmerset.cuh): the k-mers whose "$" dummies reach the graph
// Where the items of an input come from: every sequence of `seqs` yields all its items; the edges yield either all 6 per
// edge (unfiltered: k > 63, the multi-GPU driver whose ranks hold key ranges, or no room for the set) or their 2 real items
// plus 2 dummies per k-mer of the miss list.
struct ItemSource {
  const uint32_t *edges = nullptr;
  int64_t n_edges = 0;
  SeqsView seqs;
  bool filtered = false;
  const void *miss = nullptr;   // k-mers (8 bytes for k <= 31, 16 for k <= 63)
  int64_t n_miss = 0;
  int64_t n_items() const { return (filtered ? 2 * n_edges + 2 * n_miss : 6 * n_edges) + seqs.n_items; }
};
static KsGeom ks_geometry(int64_t n_edges) {
  KsGeom g;
  g.log_slots = std::max(10, ceil_log2(3.0 * (double)std::max<int64_t>(n_edges, 1)));   // >= 1.5 x the 2E entries
  g.slice_log = std::min(g.log_slots, std::max(kKsSliceLog, g.log_slots - 10));         // at most 1024 slices
  g.nslices = 1 << (g.log_slots - g.slice_log);
  return g;
}
static KsGeom ksw_geometry(int64_t n_edges) {   // k >= 64: index slots, smaller slices (the records of a slice share L2 with them)
  KsGeom g;
  g.log_slots = std::max(10, ceil_log2(3.0 * (double)std::max<int64_t>(n_edges, 1)));
  g.slice_log = std::min(g.log_slots, std::max(kKswSliceLog, g.log_slots - 10));
  g.nslices = 1 << (g.log_slots - g.slice_log);
  return g;
}
static size_t ks_bytes(int64_t n_edges, int k) {   // records + table + small tables
  if (k >= 64) return (size_t)6 * n_edges * words_tip(k) * 4 + ((size_t)8 << ksw_geometry(n_edges).log_slots) + (1 << 20);
  const size_t kb = k <= 31 ? 8 : 16;
  return (size_t)4 * n_edges * kb + ((size_t)kb << ks_geometry(n_edges).log_slots) + (1 << 20);
}
// Runs the filter on the slab (reset by the caller); the miss list is copied into c.miss so that it survives the slab.
template <int KW>
static void sdbg_filter(Ctx &c, const uint32_t *edges, int64_t n_edges, int k, ItemSource *src) {
  using Slot = typename KsKey<KW>::Slot;
  const int WK = words_key(k), WE = words_edge(k);
  const KsGeom g = ks_geometry(n_edges);
  const int nbins = 2 * g.nslices;
  Stage st(c, "items_filter");
  Slot *rec = c.alloc<Slot>((size_t)4 * n_edges);
  Slot *table = c.alloc<Slot>((size_t)1 << g.log_slots);
  unsigned long long *hist = c.alloc<unsigned long long>(nbins), *cursor = c.alloc<unsigned long long>(nbins + 3);
  MF_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * nbins, c.stream));
  MF_CUDA(cudaMemsetAsync(cursor + nbins, 0, sizeof(unsigned long long) * 3, c.stream));   // miss cursor, two tile counters
  MF_CUDA(cudaMemsetAsync(table, 0xff, sizeof(Slot) << g.log_slots, c.stream));
  const unsigned hgrid = (unsigned)std::min<int64_t>(div_ceil64(n_edges, kKsNT), (int64_t)c.sm_count * 4);
  k_ks_hist<KW><<<hgrid, kKsNT, sizeof(uint32_t) * nbins, c.stream>>>(edges, n_edges, WK, WE, k, g, hist);
  k_excl_scan_u64_small<<<1, 1024, 0, c.stream>>>(hist, nbins, cursor);
  using SC = KsScatterCfg<KW>;
  const unsigned sgrid = (unsigned)div_ceil64(n_edges, kKsNT * SC::EPT);
  const size_t ssm = ks_scatter_smem_bytes<KW>(nbins);
  if (nbins <= kKsNT) {
    set_smem(k_ks_scatter<KW, 1>, ssm);
    k_ks_scatter<KW, 1><<<sgrid, kKsNT, ssm, c.stream>>>(edges, n_edges, WK, WE, k, g, cursor, rec);
  } else {
    set_smem(k_ks_scatter<KW, 4>, ssm);
    k_ks_scatter<KW, 4><<<sgrid, kKsNT, ssm, c.stream>>>(edges, n_edges, WK, WE, k, g, cursor, rec);
  }
  // inserts are the first 2E records, queries the last 2E, both in slice order; the misses overwrite the (dead) inserts
  const unsigned wgrid = (unsigned)std::min<int64_t>(div_ceil64(2 * n_edges, kKsWalkNT * kKsWalkR), (int64_t)c.sm_count * 8);
  k_ks_insert<KW><<<wgrid, kKsWalkNT, 0, c.stream>>>(rec, 2 * n_edges, g, table, cursor + nbins + 1);
  k_ks_query<KW><<<wgrid, kKsWalkNT, 0, c.stream>>>(rec + 2 * n_edges, 2 * n_edges, g, table, rec, cursor + nbins, cursor + nbins + 2);
  MF_LAUNCH_CHECK();
  c.launches += 5;
  unsigned long long n_miss = 0;
  c.d2h(&n_miss, cursor + nbins, sizeof n_miss);
  c.miss.reserve((size_t)n_miss * sizeof(Slot) + 256);
  if (n_miss) MF_CUDA(cudaMemcpyAsync(c.miss.p, rec, (size_t)n_miss * sizeof(Slot), cudaMemcpyDeviceToDevice, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  src->filtered = true;
  src->miss = c.miss.p;
  src->n_miss = (int64_t)n_miss;
}

// k >= 64 (kmerset_w.cuh): k-mers of WR words, the table holds indices into the scattered insert records
template <int WK, int WR>
static void sdbg_filter_w(Ctx &c, const uint32_t *edges, int64_t n_edges, int k, ItemSource *src) {
  const int WE = words_edge(k);
  const KsGeom g = ksw_geometry(n_edges);
  const int nbins = 2 * g.nslices;
  Stage st(c, "items_filter");
  uint32_t *rec = c.alloc<uint32_t>((size_t)4 * n_edges * WR + 16);
  unsigned long long *table = c.alloc<unsigned long long>((size_t)1 << g.log_slots);
  unsigned long long *hist = c.alloc<unsigned long long>(nbins), *cursor = c.alloc<unsigned long long>(nbins + 3);
  MF_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * nbins, c.stream));
  MF_CUDA(cudaMemsetAsync(cursor + nbins, 0, sizeof(unsigned long long) * 3, c.stream));   // miss cursor, two tile counters
  MF_CUDA(cudaMemsetAsync(table, 0xff, sizeof(unsigned long long) << g.log_slots, c.stream));
  const unsigned hgrid = (unsigned)std::min<int64_t>(div_ceil64(n_edges, kKsNT), (int64_t)c.sm_count * 4);
  k_ksw_hist<WK, WR><<<hgrid, kKsNT, sizeof(uint32_t) * nbins, c.stream>>>(edges, n_edges, WE, k, g, hist);
  k_excl_scan_u64_small<<<1, 1024, 0, c.stream>>>(hist, nbins, cursor);
  const unsigned sgrid = (unsigned)div_ceil64(n_edges, kKsNT);
  const size_t ssm = ksw_scatter_smem_bytes<WR>(nbins);
  if (nbins <= kKsNT) {
    set_smem(k_ksw_scatter<WK, WR, 1>, ssm);
    k_ksw_scatter<WK, WR, 1><<<sgrid, kKsNT, ssm, c.stream>>>(edges, n_edges, WE, k, g, cursor, rec);
  } else {
    set_smem(k_ksw_scatter<WK, WR, 4>, ssm);
    k_ksw_scatter<WK, WR, 4><<<sgrid, kKsNT, ssm, c.stream>>>(edges, n_edges, WE, k, g, cursor, rec);
  }
  const unsigned wgrid = (unsigned)std::min<int64_t>(div_ceil64(2 * n_edges, kKsWalkNT * kKsWalkR), (int64_t)c.sm_count * 8);
  uint32_t *qrec = rec + (size_t)2 * n_edges * WR;
  // the insert records stay live during the query walk (slots point at them), so the misses get their own worst-case buffer
  uint32_t *missbuf = c.alloc<uint32_t>((size_t)2 * n_edges * WR + 16);
  k_ksw_insert<WR><<<wgrid, kKsWalkNT, 0, c.stream>>>(rec, 2 * n_edges, g, table, cursor + nbins + 1);
  k_ksw_query<WR><<<wgrid, kKsWalkNT, 0, c.stream>>>(qrec, 2 * n_edges, rec, g, table, missbuf, cursor + nbins, cursor + nbins + 2);
  MF_LAUNCH_CHECK();
  c.launches += 5;
  unsigned long long n_miss = 0;
  c.d2h(&n_miss, cursor + nbins, sizeof n_miss);
  c.miss.reserve((size_t)n_miss * WR * 4 + 256);
  if (n_miss) MF_CUDA(cudaMemcpyAsync(c.miss.p, missbuf, (size_t)n_miss * WR * 4, cudaMemcpyDeviceToDevice, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  src->filtered = true;
  src->miss = c.miss.p;
  src->n_miss = (int64_t)n_miss;
}

// MODE 0: all items of `src` into items[0 .. n_items) (returns the count); MODE 1: per-bin counts of the items' top bin_bits
// bits into hist; MODE 2: the items whose bin lies in [lo, hi) appended at *cursor.
template <int WI, int MODE>
static int64_t sdbg_generate(Ctx &c, const ItemSource &src, int k, uint32_t *items, int bin_bits, uint32_t lo, uint32_t hi,
                             unsigned long long *cursor, unsigned long long *hist) {
  if constexpr (WI >= 2) {
    const int WK = words_key(k), WE = words_edge(k);
    const size_t smem = MODE == 1 ? sizeof(uint32_t) << bin_bits : 0;
    const int64_t n_edges = src.n_edges;
    int64_t written = 0;
    // (WK, WE, WI) is one of (w,w,w), (w-1,w-1,w), (w-1,w,w) with w = WI
    if (n_edges > 0 && src.filtered) {
      const unsigned grid = (unsigned)std::min<int64_t>(div_ceil64(n_edges, kRangedNT), (int64_t)c.sm_count * 16);
      if (WK == WI) k_items_real<WI, WI, WI, MODE><<<grid, kRangedNT, smem, c.stream>>>(src.edges, n_edges, k, bin_bits, lo, hi, items, cursor, hist);
      else if (WE == WK) k_items_real<WI - 1, WI - 1, WI, MODE><<<grid, kRangedNT, smem, c.stream>>>(src.edges, n_edges, k, bin_bits, lo, hi, items, cursor, hist);
      else k_items_real<WI - 1, WI, WI, MODE><<<grid, kRangedNT, smem, c.stream>>>(src.edges, n_edges, k, bin_bits, lo, hi, items, cursor, hist);
      c.launches++;
      written = 2 * n_edges;
      if (src.n_miss > 0) {
        const unsigned mgrid = (unsigned)std::min<int64_t>(div_ceil64(src.n_miss, kRangedNT), (int64_t)c.sm_count * 16);
        uint32_t *dst = MODE == 0 ? items + (size_t)written * WI : items;
        if (k <= 31) {
          k_items_miss<1, WI, MODE><<<mgrid, kRangedNT, smem, c.stream>>>(reinterpret_cast<const unsigned long long *>(src.miss), src.n_miss, k,
                                                                         bin_bits, lo, hi, dst, cursor, hist);
        } else if (k <= 63) {
          if constexpr (WI >= 3)
            k_items_miss<2, WI, MODE><<<mgrid, kRangedNT, smem, c.stream>>>(reinterpret_cast<const unsigned __int128 *>(src.miss), src.n_miss,
                                                                           k, bin_bits, lo, hi, dst, cursor, hist);
        } else {
          // WR = words of the k-mer = WI or WI - 1
          if constexpr (WI >= 5) {
            const uint32_t *mw = reinterpret_cast<const uint32_t *>(src.miss);
            if (words_tip(k) == WI) k_items_miss_w<WI, WI, MODE><<<mgrid, kRangedNT, smem, c.stream>>>(mw, src.n_miss, k, bin_bits, lo, hi, dst, cursor, hist);
            else k_items_miss_w<WI - 1, WI, MODE><<<mgrid, kRangedNT, smem, c.stream>>>(mw, src.n_miss, k, bin_bits, lo, hi, dst, cursor, hist);
          }
        }
        c.launches++;
        written += 2 * src.n_miss;
      }
    } else if (n_edges > 0) {
      if constexpr (MODE == 0) {
        const unsigned grid = (unsigned)div_ceil64(n_edges, 128);
        if (WK == WI) k_items_from_edges<WI, WI, WI><<<grid, 128, 0, c.stream>>>(src.edges, n_edges, k, items);
        else if (WE == WK) k_items_from_edges<WI - 1, WI - 1, WI><<<grid, 128, 0, c.stream>>>(src.edges, n_edges, k, items);
        else k_items_from_edges<WI - 1, WI, WI><<<grid, 128, 0, c.stream>>>(src.edges, n_edges, k, items);
      } else {
        const unsigned grid = (unsigned)std::min<int64_t>(div_ceil64(n_edges, kRangedNT), (int64_t)c.sm_count * 8);
        constexpr bool H = MODE == 1;
        if (WK == WI) k_items_from_edges_ranged<WI, WI, WI, H><<<grid, kRangedNT, smem, c.stream>>>(src.edges, n_edges, k, bin_bits, lo, hi, items, cursor, hist);
        else if (WE == WK) k_items_from_edges_ranged<WI - 1, WI - 1, WI, H><<<grid, kRangedNT, smem, c.stream>>>(src.edges, n_edges, k, bin_bits, lo, hi, items, cursor, hist);
        else k_items_from_edges_ranged<WI - 1, WI, WI, H><<<grid, kRangedNT, smem, c.stream>>>(src.edges, n_edges, k, bin_bits, lo, hi, items, cursor, hist);
      }
      c.launches++;
      written = 6 * n_edges;
    }
    const SeqsView &sq = src.seqs;
    if (sq.n_items > 0) {
      if constexpr (MODE == 0) {
        k_items_from_seqs<WI><<<(unsigned)div_ceil64(sq.n_items, 128), 128, 0, c.stream>>>(sq.packed, sq.starts, sq.mult, sq.item_base, sq.nseq,
                                                                                        sq.n_items, k, items + (size_t)written * WI);
      } else {
        const unsigned grid = (unsigned)std::min<int64_t>(div_ceil64(sq.n_items, kRangedNT), (int64_t)c.sm_count * 8);
        k_items_from_seqs_ranged<WI, MODE == 1><<<grid, kRangedNT, smem, c.stream>>>(sq.packed, sq.starts, sq.mult, sq.item_base, sq.nseq,
                                                                                    sq.n_items, k, bin_bits, lo, hi, items, cursor, hist);
      }
      c.launches++;
      written += sq.n_items;
    }
    MF_LAUNCH_CHECK();
    return written;
  } else {
    return 0;
  }
}
static void sdbg_empty(Ctx &c, int k, SdbgView *out) {
  out->k = k;
  out->words_tip = words_tip(k);
  out->n_items = out->n_tips = out->n_large = 0;
  out->bucket_items = nullptr;
  c.sdbg_rec.reserve(256);
  c.sdbg_labels.reserve(256);
  out->rec = c.sdbg_rec.as<uint32_t>();
  out->labels = c.sdbg_labels.as<uint32_t>();
  std::fill(c.sdbg_bucket_stats.begin(), c.sdbg_bucket_stats.end(), 0);
  MF_CUDA(cudaStreamSynchronize(c.stream));
}

// items in `cur` are partitioned by their top l1_bits into the chunks of `l1` (several chunks may feed one segment);
// `other` is scratch of the same size; tables and arenas come from the slab.
template <int WI>
static void sdbg_finish(Ctx &c, uint32_t *cur, uint32_t *other, int64_t n_items, const HostChunks &l1, int l1_bits, int k,
                        int tip_mode, SdbgView *out) {
  const int Wt = words_tip(k);
  const int part_limit = 2 * (k - 1);
  out->k = k;
  out->words_tip = Wt;
  out->bucket_items = nullptr;
  c.sdbg_buckets.reserve(sizeof(unsigned long long) * kNumBuckets * 3);
  unsigned long long *d_bstats = c.sdbg_buckets.as<unsigned long long>();
  Plan p = make_plan(WI, part_limit, n_items, 1.0, false, l1_bits, l1.nseg);
  auto salloc = [&](size_t bytes) { return c.slab_alloc(bytes); };
  int bit_off = l1_bits;
  DevBuckets b = partition_chain<WI>(c, &cur, &other, l1, &bit_off, p.rest_bits, salloc, "sdbg_l2");
  int64_t *d_items = c.alloc<int64_t>(b.nslots), *d_tips = c.alloc<int64_t>(b.nslots), *d_large = c.alloc<int64_t>(b.nslots);
  int64_t *d_item_src = c.alloc<int64_t>(b.nslots), *d_tip_src = c.alloc<int64_t>(b.nslots);
  int64_t *d_item_off = c.alloc<int64_t>(b.nslots + 1), *d_tip_off = c.alloc<int64_t>(b.nslots + 1),
          *d_large_off = c.alloc<int64_t>(b.nslots + 1);
  int32_t *d_bail = c.alloc<int32_t>(b.nslots);
  int *d_flags = c.alloc<int>(4);
  unsigned long long *d_cursor = c.alloc<unsigned long long>(2);
  // arenas: emitted items <= generated items; tip labels are a small share (retry with the exact demand otherwise)
  size_t item_cap = (size_t)n_items, tip_cap = (size_t)n_items / 8 + (1 << 16);
  c.ov[0].reserve(item_cap * 4 + 256);
  DevBuf &tip_arena = c.ov[1];
  std::vector<int32_t> bail_slots;
  WorkItem *d_bail_sorted = nullptr;
  bool sorted_fallback = false;
  for (int attempt = 0;; ++attempt) {
    tip_arena.reserve(tip_cap * Wt * 4 + 256);
    MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * 4, c.stream));
    MF_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long) * 2, c.stream));
    MF_CUDA(cudaMemsetAsync(d_items, 0, sizeof(int64_t) * b.nslots, c.stream));
    MF_CUDA(cudaMemsetAsync(d_tips, 0, sizeof(int64_t) * b.nslots, c.stream));
    MF_CUDA(cudaMemsetAsync(d_large, 0, sizeof(int64_t) * b.nslots, c.stream));
    MF_CUDA(cudaMemsetAsync(d_bstats, 0, sizeof(unsigned long long) * kNumBuckets * 3, c.stream));
    LocalArgs a{};
    a.in = cur;
    a.bkt_start = b.start;
    a.bkt_size = b.size;
    a.bit_off = bit_off;
    a.sort_bits = 32 * WI - 16;   // the walker takes the largest multiplicity of equal (k-mer, b) items itself
    a.cap = p.cap;
    a.k = k;
    a.tip_mode = tip_mode;
    a.words_tip = Wt;
    a.sd_items = d_items;
    a.sd_tips = d_tips;
    a.sd_large = d_large;
    a.sd_item_off = d_item_src;
    a.sd_tip_off = d_tip_src;
    a.sd_rec = c.ov[0].as<uint32_t>();
    a.sd_labels = tip_arena.as<uint32_t>();
    a.sd_cursor = d_cursor;
    a.sd_item_cap = item_cap;
    a.sd_tip_cap = tip_cap;
    a.bucket_stats = d_bstats;
    a.bail_list = d_bail;
    a.bail_count = d_flags;
    a.overflow_flag = d_flags + 1;
    {
      Stage st(c, "local_sdbg");
      a.cap = sdbg_cap(WI);
      const size_t smem = sdbg_smem_bytes(WI, a.cap);
      set_smem(k_sdbg_local<WI>, smem);
      k_sdbg_local<WI><<<b.nslots, kSdNT, smem, c.stream>>>(a);
      MF_LAUNCH_CHECK();
      c.launches++;
    }
    int flags[2];
    c.d2h(flags, d_flags, sizeof(int) * 2);
    if (flags[0] > 0) {
      // buckets with a crowded sub-bin (low-complexity sequence) or too many items: the general LSD kernel on those slots
      Stage st(c, "local_sdbg_general");
      if (getenv("MFSDBG_TRACE")) {
        int f4[4];
        c.d2h(f4, d_flags, sizeof f4);
        fprintf(stderr, "[mfsdbg] sdbg: %d of %d buckets take the general kernel (%d too large, %d crowded; cap %d, %lld items)\n", flags[0],
                b.nslots, f4[2], f4[3], a.cap, (long long)n_items);
      }
      std::vector<int32_t> slots;
      std::vector<Range> rs = fetch_bails(c, b, d_bail, flags[0], &slots);
      std::vector<WorkItem> wi(rs.size());
      for (size_t i = 0; i < rs.size(); ++i) wi[i] = WorkItem{rs[i].start, (int32_t)std::min<int64_t>(rs[i].size, INT32_MAX), slots[i]};
      c.ov[7].reserve(sizeof(WorkItem) * wi.size());
      c.h2d(c.ov[7].p, wi.data(), sizeof(WorkItem) * wi.size());
      MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), c.stream));
      LocalArgs ga = a;
      ga.cap = local_cap(WI, false, false);
      ga.work = c.ov[7].as<WorkItem>();
      launch_local<WI, kSdbgEmit>(c, ga, (int)wi.size());
      c.d2h(flags, d_flags, sizeof(int) * 2);
      a.cap = ga.cap;
    }
    if (flags[0] > 0) {
      Stage st(c, "fallback");
      std::vector<Range> rs = fetch_bails(c, b, d_bail, flags[0], &bail_slots);
      if (!sorted_fallback) sort_ranges<WI>(c, cur, other, rs, bit_off, 32 * WI - 16);
      sorted_fallback = true;
      std::vector<WorkItem> hw;
      for (size_t i = 0; i < rs.size(); ++i) hw.push_back(WorkItem{rs[i].start, 0, bail_slots[i]});
      if (!d_bail_sorted) d_bail_sorted = c.alloc<WorkItem>(hw.size());
      c.h2d(d_bail_sorted, hw.data(), sizeof(WorkItem) * hw.size());
      a.work = d_bail_sorted;
      launch_serial<WI, kSdbgEmit>(c, a, (int)hw.size());
      c.d2h(flags, d_flags, sizeof(int) * 2);
    }
    if (!flags[1]) break;
    unsigned long long need[2];
    c.d2h(need, d_cursor, sizeof need);
    if (attempt >= 2) throw std::runtime_error("sdbg arena overflow persisted");
    tip_cap = std::max<size_t>(tip_cap, (size_t)need[1]);
  }
  Stage st(c, "sdbg_gather");
  scan_slots(c, d_items, b.nslots, d_item_off);
  scan_slots(c, d_tips, b.nslots, d_tip_off);
  scan_slots(c, d_large, b.nslots, d_large_off);
  int64_t tot[3];
  MF_CUDA(cudaMemcpyAsync(&tot[0], d_item_off + b.nslots, 8, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaMemcpyAsync(&tot[1], d_tip_off + b.nslots, 8, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaMemcpyAsync(&tot[2], d_large_off + b.nslots, 8, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  c.sdbg_rec.reserve((size_t)std::max<int64_t>(tot[0], 1) * 4 + 256);
  c.sdbg_labels.reserve((size_t)std::max<int64_t>(tot[1], 1) * Wt * 4 + 256);
  if (tot[0] > 0) {
    k_gather_edges<<<div_ceil(b.nslots, 8), 256, 0, c.stream>>>(c.ov[0].as<uint32_t>(), d_item_src, d_items, d_item_off, b.nslots, 1,
                                                              c.sdbg_rec.as<uint32_t>());
    k_gather_edges<<<div_ceil(b.nslots, 8), 256, 0, c.stream>>>(tip_arena.as<uint32_t>(), d_tip_src, d_tips, d_tip_off, b.nslots, Wt,
                                                              c.sdbg_labels.as<uint32_t>());
    MF_LAUNCH_CHECK();
    c.launches += 2;
  }
  MF_CUDA(cudaMemcpyAsync(c.sdbg_bucket_stats.data(), d_bstats, sizeof(int64_t) * kNumBuckets * 3, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  out->rec = c.sdbg_rec.as<uint32_t>();
  out->labels = c.sdbg_labels.as<uint32_t>();
  out->n_items = tot[0];
  out->n_tips = tot[1];
  out->n_large = tot[2];
}

// ---- memory-bounded sdbg: rounds over contiguous ranges of the level-1 prefix bins ------------------------------
// When the item buffers of the whole input do not fit the budget (config 5: -m 1 on error-rich reads, E ~ a third of the key
// occurrences) the stage runs like megahit's --host_mem-bounded lv1 passes: a histogram pass counts the items of every
// level-1 bin, contiguous bin ranges of at most `round_items` items are generated (ranged generators, items.cuh), sorted
// and walked one after the other, and their outputs are appended -- bin order is key order, so the concatenation is the
// global stream.  A (k-1)-prefix group never straddles a bin (l1_bits <= 11 < 2(k-1)), 16-bit buckets never do either.
template <int WI>
static void sdbg_rounds(Ctx &c, const ItemSource &src, int k, int tip_mode, int64_t round_items, size_t table_bytes, SdbgView *out) {
  const int64_t n_cap = src.n_items();
  const int Wt = words_tip(k);
  Plan p = make_plan(WI, 2 * (k - 1), n_cap, 1.0, false);
  const int nb1 = 1 << p.l1_bits;
  // per-bin item counts of the whole input
  std::vector<unsigned long long> hist(nb1);
  {
    Stage st(c, "items_hist");
    c.small[2].reserve(sizeof(unsigned long long) * ((size_t)1 << kMaxDigitBits));
    unsigned long long *d_hist = c.small[2].as<unsigned long long>();
    MF_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nb1, c.stream));
    sdbg_generate<WI, 1>(c, src, k, nullptr, p.l1_bits, 0u, (uint32_t)nb1, nullptr, d_hist);
    c.d2h(hist.data(), d_hist, sizeof(unsigned long long) * nb1);
  }
  // contiguous bin ranges of at most round_items items (a single bin may exceed it: it then is a round of its own)
  std::vector<std::pair<int, int>> rounds;
  int64_t biggest = 0;
  for (int b0 = 0; b0 < nb1;) {
    int b1 = b0;
    int64_t acc = 0;
    while (b1 < nb1 && (b1 == b0 || acc + (int64_t)hist[b1] <= round_items)) acc += (int64_t)hist[b1++];
    if (acc > 0) rounds.push_back({b0, b1});
    biggest = std::max(biggest, acc);
    b0 = b1;
  }
  if (getenv("MFSDBG_TRACE"))
    fprintf(stderr, "[mfsdbg] sdbg: %lld items in %d rounds of <= %lld (largest %lld), %d level-1 bins\n", (long long)n_cap,
            (int)rounds.size(), (long long)round_items, (long long)biggest, nb1);
  if (rounds.empty()) return sdbg_empty(c, k, out);
  c.slab_reserve((size_t)biggest * WI * 4 * 2 + 64 + table_bytes + (size_t)(biggest / 16) + (1 << 20));
  DevBuf acc_rec, acc_lab;
  int64_t n_rec = 0, n_tip = 0, n_large = 0, items_done = 0;
  std::vector<int64_t> stats((size_t)kNumBuckets * 3, 0);
  auto append = [&](DevBuf &acc, int64_t have, const uint32_t *src, int64_t add, int words, int64_t est_total) {
    const size_t need = (size_t)(have + add) * words * 4 + 256;
    if (need > acc.cap) {
      DevBuf grown;
      grown.reserve(std::max(need, (size_t)est_total * words * 4 + 256));
      if (have > 0) MF_CUDA(cudaMemcpyAsync(grown.p, acc.p, (size_t)have * words * 4, cudaMemcpyDeviceToDevice, c.stream));
      MF_CUDA(cudaStreamSynchronize(c.stream));
      acc.release();
      acc = grown;
    }
    if (add > 0)
      MF_CUDA(cudaMemcpyAsync(acc.as<uint32_t>() + (size_t)have * words, src, (size_t)add * words * 4, cudaMemcpyDeviceToDevice, c.stream));
  };
  try {
  for (const auto &rd : rounds) {
    int64_t n_r = 0;
    for (int b = rd.first; b < rd.second; ++b) n_r += (int64_t)hist[b];
    c.slab_reset();
    uint32_t *bufA = c.alloc<uint32_t>((size_t)n_r * WI + 16), *bufB = c.alloc<uint32_t>((size_t)n_r * WI + 16);
    unsigned long long *d_cur = c.alloc<unsigned long long>(1);
    {
      Stage st(c, "items");
      MF_CUDA(cudaMemsetAsync(d_cur, 0, sizeof(unsigned long long), c.stream));
      sdbg_generate<WI, 2>(c, src, k, bufA, p.l1_bits, (uint32_t)rd.first, (uint32_t)rd.second, d_cur, nullptr);
      unsigned long long got = 0;
      c.d2h(&got, d_cur, sizeof got);
      if ((int64_t)got != n_r) throw std::runtime_error("sdbg rounds: generated items disagree with the histogram");
    }
    auto salloc = [&](size_t bytes) { return c.slab_alloc(bytes); };
    HostChunks whole;
    whole.nseg = 1;
    whole.start = {0};
    whole.size = {n_r};
    whole.seg = {0};
    whole.seg_out_start = {0};
    DevBuckets b1 = partition_level<WI>(c, bufA, bufB, whole, 0, p.l1_bits, salloc, "sdbg_l1");
    std::vector<int64_t> st(nb1), sz(nb1);
    c.d2h(st.data(), b1.start, sizeof(int64_t) * nb1);
    c.d2h(sz.data(), b1.size, sizeof(int64_t) * nb1);
    HostChunks l1;
    l1.nseg = rd.second - rd.first;
    for (int i = rd.first; i < rd.second; ++i) {
      l1.start.push_back(st[i]);
      l1.size.push_back(sz[i]);
      l1.seg.push_back(i - rd.first);
      l1.seg_out_start.push_back(st[i]);
    }
    SdbgView g;
    sdbg_finish<WI>(c, bufB, bufA, n_r, l1, p.l1_bits, k, tip_mode, &g);
    items_done += n_r;
    // size the accumulators once, from the first round's yield
    const double scale = 1.1 * (double)n_cap / (double)std::max<int64_t>(items_done, 1);
    append(acc_rec, n_rec, g.rec, g.n_items, 1, (int64_t)((double)(n_rec + g.n_items) * scale) + 4096);
    append(acc_lab, n_tip, g.labels, g.n_tips, Wt, (int64_t)((double)(n_tip + g.n_tips) * scale) + 4096);
    MF_CUDA(cudaStreamSynchronize(c.stream));
    n_rec += g.n_items;
    n_tip += g.n_tips;
    n_large += g.n_large;
    for (size_t i = 0; i < stats.size(); ++i) stats[i] += c.sdbg_bucket_stats[i];
  }
  } catch (...) {   // DevBuf has no destructor: a failed round must not leak the accumulated output
    acc_rec.release();
    acc_lab.release();
    throw;
  }
  c.sdbg_rec.release();
  c.sdbg_labels.release();
  c.sdbg_rec = acc_rec;
  c.sdbg_labels = acc_lab;
  if (!c.sdbg_labels.p) c.sdbg_labels.reserve(256);
  c.sdbg_bucket_stats = stats;
  out->k = k;
  out->words_tip = Wt;
  out->bucket_items = nullptr;
  out->rec = c.sdbg_rec.as<uint32_t>();
  out->labels = c.sdbg_labels.as<uint32_t>();
  out->n_items = n_rec;
  out->n_tips = n_tip;
  out->n_large = n_large;
}

template <int WI>
static void sdbg_impl(Ctx &c, const uint32_t *edges, int64_t n_edges, const SeqsView &sq, int k, int tip_mode, SdbgView *out) {
  ItemSource src;
  src.edges = edges;
  src.n_edges = n_edges;
  src.seqs = sq;
  if (src.n_items() == 0) return sdbg_empty(c, k, out);
  const size_t held = (size_t)n_edges * words_edge(k) * 4;
  const size_t budget = c.budget();
  // ---- which dummies reach the graph (k <= 63): afterwards the item count is exact
  if (env_int("MFSDBG_ITEM_FILTER", 1) != 0 && n_edges > 0 && ks_bytes(n_edges, k) + held < budget) {
    c.slab_reserve(ks_bytes(n_edges, k) + (1 << 20));
    c.slab_reset();
    if (k <= 31) sdbg_filter<1>(c, edges, n_edges, k, &src);
    else if (k <= 63) sdbg_filter<2>(c, edges, n_edges, k, &src);
    else {
      // (WK, WR) = words of the (k+1)-mer and of the k-mer: equal unless 2k is a multiple of 32
      const int WK = words_key(k), WR = words_tip(k);
#define MF_KSW(wk, wr) if (WK == wk && WR == wr) sdbg_filter_w<wk, wr>(c, edges, n_edges, k, &src);
      MF_KSW(5, 4) MF_KSW(5, 5) MF_KSW(6, 5) MF_KSW(6, 6) MF_KSW(7, 6) MF_KSW(7, 7) MF_KSW(8, 7) MF_KSW(8, 8) MF_KSW(9, 8) MF_KSW(9, 9)
      MF_KSW(10, 9) MF_KSW(10, 10)
#undef MF_KSW
    }
  }
  const int64_t n_items = src.n_items();
  const int nb1_max = 1 << kMaxDigitBits;
  const size_t tables = (size_t)(64 << 20) + (size_t)((size_t)nb1_max << kMaxDigitBits) * 96 + (size_t)(n_items / 128);
  {
    // does the whole item set fit?  per item: two partition buffers, the record arena and the gathered records, tip labels
    const size_t per_item = (size_t)WI * 8 + 8 + (size_t)words_tip(k) / 2 + 1;
    const size_t avail = budget > held + tables ? budget - held - tables : 0;
    int64_t round_items = env_int("MFSDBG_SDBG_ROUND_ITEMS", 0);
    if (round_items <= 0 && (size_t)n_items * per_item > avail) {
      // keep half a record per generated item for the accumulated output (grown if the data yields more)
      const size_t acc = (size_t)n_items * 2;
      round_items = (int64_t)((avail > acc ? avail - acc : avail / 2) / per_item * 9 / 10);
      round_items = std::max<int64_t>(round_items, 1 << 20);
    }
    if (round_items > 0) return sdbg_rounds<WI>(c, src, k, tip_mode, round_items, tables, out);
  }
  c.slab_reserve((size_t)n_items * WI * 4 * 2 + tables + (1 << 20));
  c.slab_reset();
  uint32_t *bufA = c.alloc<uint32_t>((size_t)n_items * WI + 16), *bufB = c.alloc<uint32_t>((size_t)n_items * WI + 16);
  {
    Stage st(c, "items");
    const int64_t got = sdbg_generate<WI, 0>(c, src, k, bufA, 0, 0u, 0u, nullptr, nullptr);
    if (got != n_items) throw std::runtime_error("sdbg: generated items disagree with their count");
  }
  Plan p = make_plan(WI, 2 * (k - 1), n_items, 1.0, false);
  const int nb1 = 1 << p.l1_bits;
  auto salloc = [&](size_t bytes) { return c.slab_alloc(bytes); };
  HostChunks whole;   // level 1 over the whole item array
  whole.nseg = 1;
  whole.start = {0};
  whole.size = {n_items};
  whole.seg = {0};
  whole.seg_out_start = {0};
  DevBuckets b1 = partition_level<WI>(c, bufA, bufB, whole, 0, p.l1_bits, salloc, "sdbg_l1");
  std::vector<int64_t> st(nb1), sz(nb1);
  c.d2h(st.data(), b1.start, sizeof(int64_t) * nb1);
  c.d2h(sz.data(), b1.size, sizeof(int64_t) * nb1);
  HostChunks l1;
  l1.nseg = nb1;
  for (int i = 0; i < nb1; ++i) {
    l1.start.push_back(st[i]);
    l1.size.push_back(sz[i]);
    l1.seg.push_back(i);
    l1.seg_out_start.push_back(st[i]);
  }
  sdbg_finish<WI>(c, bufB, bufA, n_items, l1, p.l1_bits, k, tip_mode, out);
}
#define MF_DISPATCH_CASE_SDBG(Wn) \
  case Wn: sdbg_impl<Wn>(c, edges, n_edges, seqs, k, tip_mode, out); break;
void dev_seq2sdbg(Ctx &c, const uint32_t *edges, int64_t n_edges, const SeqsView &seqs, int k, int tip_mode, SdbgView *out) {
  if (k < 9 || k > 150) throw std::invalid_argument("k must be in [9, 150]");
  MF_DISPATCH_W(words_item(k), SDBG)
}

// w | last<<4 | tip<<5 | min(mult, 255)<<8 per item (SdbgWriter::Write); items with mult > 254 also go to `pairs` as
// (index << 16 | mult), unordered (the host sorts the few of them)
__global__ void k_sdbg_pack16(const uint32_t *__restrict__ rec, int64_t n, uint16_t *__restrict__ out, unsigned long long *pairs,
                              unsigned long long *cursor) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = rec[i], m = r >> 8;
    out[i] = (uint16_t)((r & 0x3fu) | (min(m, 255u) << 8));
    if (m > 254u) pairs[atomicAdd(cursor, 1ull)] = ((unsigned long long)i << 16) | (unsigned long long)min(m, 65535u);
  }
}
void dev_sdbg_pack16(Ctx &c, const SdbgView &g, uint16_t *rec16_dev, unsigned long long *pairs_dev, unsigned long long *cursor_dev) {
  MF_CUDA(cudaMemsetAsync(cursor_dev, 0, sizeof(unsigned long long), c.stream));
  if (g.n_items > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>(div_ceil64(g.n_items, 256), (int64_t)c.sm_count * 16);
    k_sdbg_pack16<<<grid, 256, 0, c.stream>>>(g.rec, g.n_items, rec16_dev, pairs_dev, cursor_dev);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
}

// ---- staged sdbg for the multi-GPU driver: items -> (caller exchanges them by prefix) -> finish
#define MF_DISPATCH_CASE_SITEMS(Wn)                                                  \
  case Wn: {                                                                          \
    ItemSource src;                                                                   \
    src.edges = edges;                                                                \
    src.n_edges = n_edges;                                                            \
    Stage st(c, "items");                                                             \
    sdbg_generate<Wn, 0>(c, src, k, items_out, 0, 0u, 0u, nullptr, nullptr);          \
  } break;
void dev_sdbg_items(Ctx &c, const uint32_t *edges, int64_t n_edges, int k, uint32_t *items_out) {
  if (k < 9 || k > 150) throw std::invalid_argument("k must be in [9, 150]");
  MF_DISPATCH_W(words_item(k), SITEMS)
}
// the same plus every item of the sequences (contigs): 6 n_edges + seqs.n_items items
#define MF_DISPATCH_CASE_SITEMSQ(Wn)                                                 \
  case Wn: {                                                                          \
    ItemSource src;                                                                   \
    src.edges = edges;                                                                \
    src.n_edges = n_edges;                                                            \
    src.seqs = seqs;                                                                  \
    Stage st(c, "items");                                                             \
    return sdbg_generate<Wn, 0>(c, src, k, items_out, 0, 0u, 0u, nullptr, nullptr);   \
  }
int64_t dev_sdbg_items_seqs(Ctx &c, const uint32_t *edges, int64_t n_edges, const SeqsView &seqs, int k, uint32_t *items_out) {
  if (k < 9 || k > 150) throw std::invalid_argument("k must be in [9, 150]");
  MF_DISPATCH_W(words_item(k), SITEMSQ)
  return 0;
}
// generic: histogram / partition of W-word records by their top l1_bits (one segment)
template <int W>
static void records_hist_impl(Ctx &c, const uint32_t *rec, int64_t n, int l1_bits, unsigned long long *hist_dev) {
  using C = TileCfg<W>;
  const int nb = 1 << l1_bits;
  MF_CUDA(cudaMemsetAsync(hist_dev, 0, sizeof(unsigned long long) * nb, c.stream));
  const int64_t tiles = div_ceil64(n, C::TH);
  if (tiles == 0) return;
  c.ov[2].reserve(sizeof(TileDesc) * tiles + 64);
  c.ov[3].reserve(64);
  int64_t h[4] = {0, n, 0, tiles};   // start, size, tile_base[0], tile_base[1]
  int32_t seg0 = 0;
  c.h2d(c.ov[3].p, h, sizeof h);
  c.h2d((char *)c.ov[3].p + 40, &seg0, 4);
  const int64_t *dp = c.ov[3].as<int64_t>();
  ChunkTable ct{dp, dp + 1, reinterpret_cast<const int32_t *>((const char *)c.ov[3].p + 40), dp + 2, 1};
  k_build_tiles<<<(unsigned)div_ceil64(tiles, 256), 256, 0, c.stream>>>(ct, C::TH, tiles, c.ov[2].as<TileDesc>());
  RecordsProducer<W> ph{rec, c.ov[2].as<TileDesc>(), C::TH};
  size_t smem = ((size_t)1 << l1_bits) * 4 + 16;
  auto kern = k_level_hist<RecordsProducer<W>, W, C::NT, C::IPT_H>;
  set_smem(kern, smem);
  Stage st(c, "records_hist");
  kern<<<(unsigned)tiles, C::NT, smem, c.stream>>>(ph, LevelArgs{0, l1_bits, 0u, (uint32_t)nb, nullptr}, hist_dev);
  MF_LAUNCH_CHECK();
  c.launches += 2;
}
template <int W>
static void records_scatter_impl(Ctx &c, const uint32_t *rec, int64_t n, int l1_bits, const unsigned long long *hist_dev, uint32_t *out,
                                 const unsigned long long *bin_base) {
  using C = TileCfg<W>;
  const int nb = 1 << l1_bits;
  const int64_t tiles = div_ceil64(n, C::TS);
  if (tiles == 0) return;
  c.ov[2].reserve(sizeof(TileDesc) * tiles + 64);
  c.ov[3].reserve(64);
  c.ov[4].reserve(sizeof(unsigned long long) * kMaxBins);
  int64_t h[4] = {0, n, 0, tiles};
  int32_t seg0 = 0;
  c.h2d(c.ov[3].p, h, sizeof h);
  c.h2d((char *)c.ov[3].p + 40, &seg0, 4);
  const int64_t *dp = c.ov[3].as<int64_t>();
  ChunkTable ct{dp, dp + 1, reinterpret_cast<const int32_t *>((const char *)c.ov[3].p + 40), dp + 2, 1};
  k_build_tiles<<<(unsigned)div_ceil64(tiles, 256), 256, 0, c.stream>>>(ct, C::TS, tiles, c.ov[2].as<TileDesc>());
  if (bin_base) MF_CUDA(cudaMemsetAsync(c.ov[4].p, 0, sizeof(unsigned long long) * nb, c.stream));
  else k_excl_scan_u64_small<<<1, 1024, 0, c.stream>>>(hist_dev, nb, c.ov[4].as<unsigned long long>());
  RecordsProducer<W> ps{rec, c.ov[2].as<TileDesc>(), C::TS};
  size_t smem = scatter_smem_bytes<W>(C::NT, C::TS, l1_bits, 4);
  auto kern = level_scatter_kernel<W>(l1_bits);
  set_smem(kern, smem);
  Stage st(c, "records_scatter");
  kern<<<(unsigned)tiles, C::NT, smem, c.stream>>>(ps, LevelArgs{0, l1_bits, 0u, (uint32_t)nb, bin_base}, c.ov[4].as<unsigned long long>(), out);
  MF_LAUNCH_CHECK();
  c.launches += 3;
}
#define MF_DISPATCH_CASE_RHIST(Wn) \
  case Wn: records_hist_impl<Wn>(c, rec, n, l1_bits, hist_dev); break;
void dev_records_hist(Ctx &c, const uint32_t *rec, int64_t n, int words, int l1_bits, unsigned long long *hist_dev) {
  if (l1_bits < 1 || l1_bits > kMaxDigitBits) throw std::invalid_argument("l1_bits must be in [1, 11]");
  MF_DISPATCH_W(words, RHIST)
}
#define MF_DISPATCH_CASE_RSCAT(Wn) \
  case Wn: records_scatter_impl<Wn>(c, rec, n, l1_bits, hist_dev, out, bin_base); break;
void dev_records_scatter(Ctx &c, const uint32_t *rec, int64_t n, int words, int l1_bits, const unsigned long long *hist_dev,
                         uint32_t *out, const unsigned long long *bin_base) {
  if (l1_bits < 1 || l1_bits > kMaxDigitBits) throw std::invalid_argument("l1_bits must be in [1, 11]");
  MF_DISPATCH_W(words, RSCAT)
}
template <int WI>
static void dev_sdbg_finish_w(Ctx &c, uint32_t *items, uint32_t *scratch, int64_t n_items, const HostChunks &hc, int k, int l1_bits,
                              int tip_mode, SdbgView *out) {
  if (n_items == 0) return sdbg_empty(c, k, out);
  const size_t table_bytes = (size_t)(64 << 20) + (size_t)((size_t)hc.nseg << kMaxDigitBits) * 96 + (size_t)(n_items / 16);
  c.slab_reserve(table_bytes + (1 << 20));
  sdbg_finish<WI>(c, items, scratch, n_items, hc, l1_bits, k, tip_mode, out);
}
#define MF_DISPATCH_CASE_SFIN(Wn) \
  case Wn: dev_sdbg_finish_w<Wn>(c, items, scratch, n_items, hc, k, l1_bits, tip_mode, out); break;
void dev_sdbg_finish(Ctx &c, uint32_t *items, uint32_t *scratch, int64_t n_items, const int64_t *chunk_start,
                     const int64_t *chunk_size, const int32_t *chunk_seg, int n_chunks, int n_segs, int k, int l1_bits, int tip_mode,
                     SdbgView *out) {
  HostChunks hc;
  hc.nseg = n_segs;
  hc.start.assign(chunk_start, chunk_start + n_chunks);
  hc.size.assign(chunk_size, chunk_size + n_chunks);
  hc.seg.assign(chunk_seg, chunk_seg + n_chunks);
  std::vector<int64_t> tot(n_segs, 0);
  for (int i = 0; i < n_chunks; ++i) {
    if (chunk_seg[i] < 0 || chunk_seg[i] >= n_segs) throw std::invalid_argument("chunk_seg out of range");
    tot[chunk_seg[i]] += chunk_size[i];
  }
  int64_t acc = 0;
  for (int s2 = 0; s2 < n_segs; ++s2) { hc.seg_out_start.push_back(acc); acc += tot[s2]; }
  if (acc != n_items) throw std::invalid_argument("chunk sizes do not add up to n_items");
  MF_DISPATCH_W(words_item(k), SFIN)
}

}  // namespace mf
