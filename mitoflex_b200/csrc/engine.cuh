// engine.cuh -- host-side orchestration of the device pipelines (one context per GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace mf {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes);
  void release();
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {   // pinned host memory
  void *p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes);
  void release();
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

struct StageRec {
  std::string name;
  cudaEvent_t e0, e1;
  double host_ms = 0;   // host clock at stage_begin, relative to begin_call (MFSDBG_TRACE)
};

struct Ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;        // host-input pipeline: chunked H2D runs ahead of the reads-fed kernels
  std::vector<cudaEvent_t> copy_events;
  char *slab = nullptr;
  size_t slab_bytes = 0, slab_off = 0;
  size_t mem_limit = 0;
  int64_t launches = 0;
  bool profiling = false;
  std::vector<StageRec> stages;
  std::vector<int> open_stages;
  std::string profile;
  double call_t0 = 0;
  // results that outlive a call
  DevBuf edges, sdbg_rec, sdbg_labels, sdbg_buckets, sbits, pack_words, pack_starts, synth_words, synth_starts, in_words, in_starts;
  HostBuf out_rec, out_labels, out_large;
  HostBuf io_pin[2];                          // file writers: two pinned pieces of the output stream (hostio.cu)
  std::vector<int64_t> out_large_index;
  std::vector<uint16_t> out_large_mult;
  DevBuf small[3];   // per-call histograms / counters kept across calls (cudaMalloc and cudaFree stall for 100+ ms at times)
  DevBuf fb[4];      // scratch of the recursive fallback sort, one per recursion depth
  DevBuf miss;     // miss list of the sdbg item filter (kmerset.cuh): outlives the slab across the rounds of a call
  DevBuf ov[10];   // scratch of the oversized-bucket path, kept across calls (cudaMalloc/cudaFree are slow and synchronise)
  // tables read back by the file-level API
  std::vector<int64_t> edge_bucket_counts;   // 65536
  std::vector<int64_t> sdbg_bucket_stats;    // 65536 * 3 (items, tips, large)

  explicit Ctx(int dev);
  ~Ctx();
  size_t budget();                 // bytes this context may hold (slab + results)
  void slab_reserve(size_t bytes); // (re)allocate the slab; invalidates earlier slab pointers
  void slab_reset() { slab_off = 0; }
  void *slab_alloc(size_t bytes);
  template <class T>
  T *alloc(size_t n) { return reinterpret_cast<T *>(slab_alloc(n * sizeof(T))); }
  // synchronous copies ordered on this context's stream (never the legacy default stream)
  void d2h(void *dst, const void *src, size_t bytes);
  void h2d(void *dst, const void *src, size_t bytes);
  void begin_call();
  void end_call();
  void stage_begin(const char *name);
  void stage_end();
};

struct Stage {
  Ctx &c;
  Stage(Ctx &ctx, const char *name) : c(ctx) { c.stage_begin(name); }
  ~Stage() { c.stage_end(); }
};

struct ReadsView {
  const uint32_t *packed;
  const int64_t *starts;
  int64_t n_reads, n_bases;
};
struct EdgesView {
  const uint32_t *edges = nullptr;
  int64_t n_edges = 0, n_keys = 0;
  int k = 0, words = 0;
};
struct SdbgView {
  const uint32_t *rec = nullptr, *labels = nullptr;
  const int64_t *bucket_items = nullptr;
  int64_t n_items = 0, n_tips = 0, n_large = 0;
  int k = 0, words_tip = 0;
};
struct SeqsView {          // general sequences (contigs), stored orientation, on device
  const uint32_t *packed = nullptr;
  const int64_t *starts = nullptr;   // device, nseq+1
  const uint16_t *mult = nullptr;    // device
  const int64_t *item_base = nullptr;  // device, nseq+1
  int nseq = 0;
  int64_t n_items = 0;
};

// pipelines
void dev_count(Ctx &c, const ReadsView &r, int k, int min_count, EdgesView *out, int64_t *counting_host);
// count with the reads still in (pinned) host memory: the transfer is cut into chunks and the reads-fed partition level runs
// on chunk i while chunk i+1 crosses PCIe.  Returns false (nothing done) when the job does not fit one in-core round.
bool dev_count_host(Ctx &c, const uint32_t *packed_host, const int64_t *starts_host, int64_t n_reads, int64_t n_bases, int k,
                    int min_count, EdgesView *out);
void dev_seq2sdbg(Ctx &c, const uint32_t *edges, int64_t n_edges, const SeqsView &seqs, int k, int tip_mode, SdbgView *out);
// the graph's records as megahit's 16-bit packed items + (index, multiplicity) pairs of the items beyond 254, for the trip over PCIe
void dev_sdbg_pack16(Ctx &c, const SdbgView &g, uint16_t *rec16_dev, unsigned long long *pairs_dev, unsigned long long *cursor_dev);
void dev_count_hist(Ctx &c, const ReadsView &r, int k, int l1_bits, unsigned long long *hist_dev);
// capacity: records keys_out can hold (checked against the histogram total; < 0 = unchecked, peer mode ignores it)
void dev_count_scatter(Ctx &c, const ReadsView &r, int k, int l1_bits, const unsigned long long *hist_dev, uint32_t *keys_out,
                       int64_t capacity, const unsigned long long *bin_base);
void dev_count_finish(Ctx &c, uint32_t *keys, uint32_t *scratch, int64_t n_keys, const int64_t *chunk_start,
                      const int64_t *chunk_size, const int32_t *chunk_seg, int n_chunks, int n_segs, int k, int l1_bits,
                      int min_count, EdgesView *out, int64_t *counting_host);

// the read-start bitmap of `r` (1 bit per base), rebuilt on every call into the context's buffer
const uint32_t *dev_start_bits(Ctx &c, const ReadsView &r);

// ---- super-k-mer exchange (skm.cu): the multi-GPU count sends runs of consecutive (k+1)-mers that share an owner as one
// 64-bit record (2-bit bases + run length) instead of one 8-byte key per (k+1)-mer
bool skm_supported(int k);
int64_t skm_key_capacity(int64_t n_keys);
// dst_ptrs == nullptr: count only, on every `stride`-th tile.  counts_out[2 * n_dst]: records, then keys, per destination
void dev_skm_scatter(Ctx &c, const ReadsView &r, int k, int n_dst, const uint64_t *dst_ptrs, const int64_t *dst_caps, int64_t stride,
                     int64_t *counts_out);
void dev_count_skm(Ctx &c, const uint64_t *recs, const int64_t *chunk_start, const int64_t *chunk_size, int n_chunks, int64_t n_keys,
                   int k, int min_count, uint32_t *keys, uint32_t *scratch, int64_t capacity, EdgesView *out);

// ---- the item filter across GPUs (ksdist.cu): hash slices of the k-mer set dealt out to the GPUs, records stored into the owner's
// buffer by the scatter kernel, every owner probes its slices and keeps its share of the miss list in the context
bool ksd_supported(int k);
void ksd_geometry(int64_t n_edges_global, int world, int *log_slots, int *slice_log);
void dev_ksd_hist(Ctx &c, const uint32_t *edges, int64_t n_edges, int k, int log_slots, int slice_log, unsigned long long *hist_dev);
void dev_ksd_scatter(Ctx &c, const uint32_t *edges, int64_t n_edges, int k, int log_slots, int slice_log,
                     const unsigned long long *bin_base_dev);
int64_t dev_ksd_filter(Ctx &c, const uint64_t *ins, int64_t n_ins, const uint64_t *qry, int64_t n_qry, int log_slots, int slice_log,
                       int slice_lo, int n_owned);
int64_t dev_ksd_items(Ctx &c, const uint32_t *edges, int64_t n_edges, int64_t n_miss, int k, uint32_t *items_out, int64_t capacity);

void dev_sdbg_items(Ctx &c, const uint32_t *edges, int64_t n_edges, int k, uint32_t *items_out);
int64_t dev_sdbg_items_seqs(Ctx &c, const uint32_t *edges, int64_t n_edges, const SeqsView &seqs, int k, uint32_t *items_out);
void dev_records_hist(Ctx &c, const uint32_t *rec, int64_t n, int words, int l1_bits, unsigned long long *hist_dev);
void dev_records_scatter(Ctx &c, const uint32_t *rec, int64_t n, int words, int l1_bits, const unsigned long long *hist_dev,
                         uint32_t *out, const unsigned long long *bin_base);
void dev_sdbg_finish(Ctx &c, uint32_t *items, uint32_t *scratch, int64_t n_items, const int64_t *chunk_start,
                     const int64_t *chunk_size, const int32_t *chunk_seg, int n_chunks, int n_segs, int k, int l1_bits, int tip_mode,
                     SdbgView *out);

}  // namespace mf
