// ksdist.cu -- the item filter of seq2sdbg (kmerset.cuh) across GPUs, 16 <= k <= 31.
//
// Which "$" dummies of an edge set reach the graph is a membership question about k-mers: a suffix k-mer of some edge that is
// the prefix k-mer of no edge (of the edge set closed under reverse complement) is a dead end and gets its two dummies; every
// other dummy dies in Lv2Postprocess (megahit sdbg/sdbg_builder / SeqToSdbg).  With the edges split over GPUs -- by key range
// or, after the super-k-mer count, pseudo-randomly -- no GPU can answer that alone, and the first multi-GPU driver generated
// all 6 items per edge (3x the items through the item exchange and the sort: local_sdbg 23 ms against 7).  Here the hash
// slices of the k-mer set are dealt out to the GPUs; every GPU cuts its edges into the same 4 k-mer records per edge the
// single-GPU filter makes, and the scatter kernel (k_ks_scatter, bin_base set) stores each record into the buffer of the GPU
// that owns its slice -- over NVLink peer memory, slice-major so that the owner's insert / query walks keep their L2 window.
// The owner builds its part of the table, probes it, and ends up with its share of the miss list; the driver then generates
// 2 real items per local edge plus 2 dummies per local miss and exchanges items by prefix as before.
#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include "engine.cuh"
#include "kmerset.cuh"

namespace mf {

bool ksd_supported(int k) { return k >= 16 && k <= 31; }

static int ksd_ceil_log2(double x) {
  int b = 0;
  while ((double)(1ull << b) < x && b < 62) ++b;
  return b;
}
// table geometry for `n_edges_global` edges over `world` GPUs: as ks_geometry (engine.cu), but cut into at least `world` slices
void ksd_geometry(int64_t n_edges_global, int world, int *log_slots, int *slice_log) {
  const int ls = std::max(10, ksd_ceil_log2(3.0 * (double)std::max<int64_t>(n_edges_global, 1)));
  int min_sl = kKsSliceLog;                                        // 16 MB slices; MFSDBG_KS_SLICE_LOG (21..24) asks for larger ones:
  if (const char *e = getenv("MFSDBG_KS_SLICE_LOG"))               // fewer exchange bins = longer NVLink runs, weaker L2 window
    if (*e) min_sl = std::max(kKsSliceLog, std::min(24, atoi(e)));
  int sl = std::min(ls, std::max(min_sl, ls - 10));               // at most 1024 slices of >= 16 MB
  const int wl = ksd_ceil_log2((double)std::max(world, 1));
  if (ls - sl < wl) sl = std::max(6, ls - wl);                    // small inputs: still a slice per GPU
  *log_slots = ls;
  *slice_log = sl;
}
static KsGeom ksd_geom(int log_slots, int slice_log) {
  if (log_slots < 6 || log_slots > 40 || slice_log < 6 || slice_log > log_slots || log_slots - slice_log > 10)
    throw std::invalid_argument("k-mer set geometry out of range");
  KsGeom g;
  g.log_slots = log_slots;
  g.slice_log = slice_log;
  g.nslices = 1 << (log_slots - slice_log);
  return g;
}
template <class K>
static void ksd_set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) throw std::runtime_error("kernel shared memory request exceeds 227 KB");
  MF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// records per (kind, slice) of this GPU's edges: hist_dev[kind * nslices + slice], kind 0 = inserts, 1 = queries
void dev_ksd_hist(Ctx &c, const uint32_t *edges, int64_t n_edges, int k, int log_slots, int slice_log, unsigned long long *hist_dev) {
  if (!ksd_supported(k)) throw std::invalid_argument("the multi-GPU item filter needs 16 <= k <= 31");
  const KsGeom g = ksd_geom(log_slots, slice_log);
  const int nbins = 2 * g.nslices;
  MF_CUDA(cudaMemsetAsync(hist_dev, 0, sizeof(unsigned long long) * nbins, c.stream));
  if (n_edges == 0) return;
  Stage st(c, "ks_hist");
  const unsigned grid = (unsigned)std::min<int64_t>(div_ceil64(n_edges, kKsNT), (int64_t)c.sm_count * 4);
  k_ks_hist<1><<<grid, kKsNT, sizeof(uint32_t) * nbins, c.stream>>>(edges, n_edges, words_key(k), words_edge(k), k, g, hist_dev);
  MF_LAUNCH_CHECK();
  c.launches++;
}

// the 4 k-mer records of every edge stored at bin_base_dev[kind * nslices + slice] + (running count of that bin) * 8
void dev_ksd_scatter(Ctx &c, const uint32_t *edges, int64_t n_edges, int k, int log_slots, int slice_log,
                     const unsigned long long *bin_base_dev) {
  if (!ksd_supported(k)) throw std::invalid_argument("the multi-GPU item filter needs 16 <= k <= 31");
  const KsGeom g = ksd_geom(log_slots, slice_log);
  const int nbins = 2 * g.nslices;
  if (n_edges == 0) return;
  c.slab_reserve(1 << 20);
  unsigned long long *cursor = c.alloc<unsigned long long>(nbins);
  MF_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long) * nbins, c.stream));
  Stage st(c, "ks_scatter");
  using SC = KsScatterCfg<1>;
  const unsigned grid = (unsigned)div_ceil64(n_edges, kKsNT * SC::EPT);
  const size_t smem = ks_scatter_smem_bytes<1>(nbins);
  if (nbins <= kKsNT) {
    ksd_set_smem(k_ks_scatter<1, 1>, smem);
    k_ks_scatter<1, 1><<<grid, kKsNT, smem, c.stream>>>(edges, n_edges, words_key(k), words_edge(k), k, g, cursor, nullptr, bin_base_dev);
  } else {
    ksd_set_smem(k_ks_scatter<1, 4>, smem);
    k_ks_scatter<1, 4><<<grid, kKsNT, smem, c.stream>>>(edges, n_edges, words_key(k), words_edge(k), k, g, cursor, nullptr, bin_base_dev);
  }
  MF_LAUNCH_CHECK();
  c.launches++;
}

// this GPU's slices [slice_lo, slice_lo + n_owned): build the table from the received inserts, probe it with the received
// queries (both slice-major); the misses end up in the context's miss list.  Returns their number.
int64_t dev_ksd_filter(Ctx &c, const uint64_t *ins, int64_t n_ins, const uint64_t *qry, int64_t n_qry, int log_slots, int slice_log,
                       int slice_lo, int n_owned) {
  const KsGeom g = ksd_geom(log_slots, slice_log);
  if (slice_lo < 0 || n_owned < 0 || slice_lo + n_owned > g.nslices) throw std::invalid_argument("slice range outside the table");
  if (n_qry == 0 || n_owned == 0) return 0;
  using Slot = unsigned long long;
  const size_t slots = (size_t)n_owned << g.slice_log;
  c.slab_reserve(slots * sizeof(Slot) + (size_t)n_qry * sizeof(Slot) + (1 << 20));
  Slot *table = c.alloc<Slot>(slots);
  Slot *miss = c.alloc<Slot>((size_t)n_qry);
  unsigned long long *ctr = c.alloc<unsigned long long>(4);   // miss cursor, two tile counters
  MF_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long) * 4, c.stream));
  MF_CUDA(cudaMemsetAsync(table, 0xff, slots * sizeof(Slot), c.stream));
  Slot *tab0 = table - ((size_t)slice_lo << g.slice_log);     // the kernels index the table by GLOBAL slot
  unsigned long long n_miss = 0;
  {
    Stage st(c, "items_filter");
    const unsigned wgrid = (unsigned)std::min<int64_t>(div_ceil64(std::max(n_ins, n_qry), kKsWalkNT * kKsWalkR), (int64_t)c.sm_count * 8);
    if (n_ins > 0) k_ks_insert<1><<<wgrid, kKsWalkNT, 0, c.stream>>>(reinterpret_cast<const Slot *>(ins), n_ins, g, tab0, ctr + 1);
    k_ks_query<1><<<wgrid, kKsWalkNT, 0, c.stream>>>(reinterpret_cast<const Slot *>(qry), n_qry, g, tab0, miss, ctr, ctr + 2);
    MF_LAUNCH_CHECK();
    c.launches += 2;
  }
  c.d2h(&n_miss, ctr, sizeof n_miss);
  c.miss.reserve((size_t)n_miss * sizeof(Slot) + 256);
  if (n_miss) MF_CUDA(cudaMemcpyAsync(c.miss.p, miss, (size_t)n_miss * sizeof(Slot), cudaMemcpyDeviceToDevice, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  return (int64_t)n_miss;
}

// 2 real items per edge, then 2 dummies per k-mer of the miss list dev_ksd_filter left in the context
template <int WK, int WE, int WI>
static void ksd_items_w(Ctx &c, const uint32_t *edges, int64_t n_edges, int64_t n_miss, int k, uint32_t *items) {
  if (n_edges > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>(div_ceil64(n_edges, kRangedNT), (int64_t)c.sm_count * 16);
    k_items_real<WK, WE, WI, 0><<<grid, kRangedNT, 0, c.stream>>>(edges, n_edges, k, 0, 0u, 0u, items, nullptr, nullptr);
    c.launches++;
  }
  if (n_miss > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>(div_ceil64(n_miss, kRangedNT), (int64_t)c.sm_count * 16);
    k_items_miss<1, WI, 0><<<grid, kRangedNT, 0, c.stream>>>(c.miss.as<unsigned long long>(), n_miss, k, 0, 0u, 0u,
                                                             items + (size_t)2 * n_edges * WI, nullptr, nullptr);
    c.launches++;
  }
  MF_LAUNCH_CHECK();
}
int64_t dev_ksd_items(Ctx &c, const uint32_t *edges, int64_t n_edges, int64_t n_miss, int k, uint32_t *items_out, int64_t capacity) {
  if (!ksd_supported(k)) throw std::invalid_argument("the multi-GPU item filter needs 16 <= k <= 31");
  const int64_t n = 2 * n_edges + 2 * n_miss;
  if (n > capacity) throw std::invalid_argument("items_out holds " + std::to_string(capacity) + " items, " + std::to_string(n) + " needed");
  if ((size_t)n_miss * 8 > c.miss.cap) throw std::invalid_argument("n_miss exceeds the miss list held by the context");
  Stage st(c, "items");
  const int WK = words_key(k), WE = words_edge(k), WI = words_item(k);
  if (WK == 2 && WE == 2 && WI == 2) ksd_items_w<2, 2, 2>(c, edges, n_edges, n_miss, k, items_out);
  else if (WK == 2 && WE == 2 && WI == 3) ksd_items_w<2, 2, 3>(c, edges, n_edges, n_miss, k, items_out);
  else if (WK == 2 && WE == 3 && WI == 3) ksd_items_w<2, 3, 3>(c, edges, n_edges, n_miss, k, items_out);
  else throw std::runtime_error("unexpected record widths");
  return n;
}

}  // namespace mf
