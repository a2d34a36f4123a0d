// api.cu -- the C ABI of libmfsdbg.so (include/mfsdbg.h): argument checking, error convention, no exceptions out.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "engine.cuh"
#include "hostio.h"
#include "mfsdbg.h"

namespace mf {
void dev_synth(Ctx &c, const mfsdbg_synth_spec &sp, ReadsView *out);
void dev_pack_fastq(Ctx &c, const uint8_t *const *texts, const int64_t *n_bytes, int n_texts, int n_policy, ReadsView *out,
                    int *max_len);
}  // namespace mf

struct mfsdbg_ctx {
  mf::Ctx c;
  explicit mfsdbg_ctx(int dev) : c(dev) {}
};

static thread_local std::string g_err;
static std::mutex g_job_mutex;   // one job owns the GPUs at a time (SURVEY.md 8b threading)

const char *mfsdbg_last_error(void) { return g_err.c_str(); }
int mfsdbg_version(void) { return MFSDBG_VERSION; }
int mfsdbg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <class F>
static int guarded(F &&f) {
  try {
    g_err.clear();
    f();
    return MFSDBG_OK;
  } catch (const mf::CudaError &e) {
    g_err = e.what();
    cudaGetLastError();
    return MFSDBG_ECUDA;
  } catch (const mf::IoError &e) {
    g_err = e.what();
    return MFSDBG_EIO;
  } catch (const std::invalid_argument &e) {
    g_err = e.what();
    return MFSDBG_EINVAL;
  } catch (const std::bad_alloc &) {
    g_err = "host memory exhausted";
    return MFSDBG_ENOMEM;
  } catch (const std::exception &e) {
    g_err = e.what();
    return MFSDBG_EINTERNAL;
  } catch (...) {
    g_err = "unknown failure";
    return MFSDBG_EINTERNAL;
  }
}

mfsdbg_ctx *mfsdbg_ctx_create(int32_t device) {
  if (device < 0 || device >= mfsdbg_device_count()) {
    g_err = "no usable CUDA device " + std::to_string(device) + " (libmfsdbg has no CPU fallback)";
    return nullptr;
  }
  mfsdbg_ctx *ctx = nullptr;
  int rc = guarded([&] { ctx = new mfsdbg_ctx(device); });
  return rc == MFSDBG_OK ? ctx : nullptr;
}
void mfsdbg_ctx_destroy(mfsdbg_ctx *ctx) { delete ctx; }
int mfsdbg_ctx_set_mem_limit(mfsdbg_ctx *ctx, uint64_t bytes) {
  if (!ctx) return MFSDBG_EINVAL;
  ctx->c.mem_limit = (size_t)bytes;
  return MFSDBG_OK;
}
int mfsdbg_ctx_set_stream(mfsdbg_ctx *ctx, void *cuda_stream) {
  if (!ctx) return MFSDBG_EINVAL;
  cudaStreamSynchronize(ctx->c.stream);
  ctx->c.stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->c.own_stream;
  return MFSDBG_OK;
}
void *mfsdbg_ctx_stream(mfsdbg_ctx *ctx) { return ctx ? (void *)ctx->c.stream : nullptr; }
int64_t mfsdbg_ctx_launches(mfsdbg_ctx *ctx) { return ctx ? ctx->c.launches : 0; }
int mfsdbg_ctx_set_profiling(mfsdbg_ctx *ctx, int32_t on) {
  if (!ctx) return MFSDBG_EINVAL;
  ctx->c.profiling = on != 0;
  return MFSDBG_OK;
}
const char *mfsdbg_ctx_last_profile(mfsdbg_ctx *ctx) { return ctx ? ctx->c.profile.c_str() : ""; }
int32_t mfsdbg_words_per_key(int32_t k) { return mf::words_key(k); }
int32_t mfsdbg_words_per_edge(int32_t k) { return mf::words_edge(k); }

static mf::ReadsView view(const mfsdbg_dev_reads *r) {
  if (!r || r->n_reads < 0 || r->n_bases < 0 || (r->n_bases > 0 && (!r->packed || !r->starts)))
    throw std::invalid_argument("bad mfsdbg_dev_reads");
  return mf::ReadsView{r->packed, r->starts, r->n_reads, r->n_bases};
}
static void fill(mfsdbg_dev_edges *o, const mf::EdgesView &e) {
  o->edges = e.edges;
  o->n_edges = e.n_edges;
  o->k = e.k;
  o->words_per_edge = e.words;
  o->n_keys = e.n_keys;
}
static void fill(mfsdbg_dev_sdbg *o, const mf::SdbgView &g, mf::Ctx &c) {
  o->rec = g.rec;
  o->tip_labels = g.labels;
  o->bucket_items = nullptr;
  o->n_items = g.n_items;
  o->n_tips = g.n_tips;
  o->n_large = g.n_large;
  o->k = g.k;
  o->words_per_tip = g.words_tip;
  (void)c;
}

int mfsdbg_dev_count(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t min_count, mfsdbg_dev_edges *out,
                     int64_t *counting_host) {
  if (!ctx || !out) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::EdgesView e;
    mf::dev_count(ctx->c, view(reads), k, min_count, &e, counting_host);
    ctx->c.end_call();
    fill(out, e);
  });
}
int mfsdbg_dev_seq2sdbg(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int32_t k, int32_t tip_mode,
                        mfsdbg_dev_sdbg *out) {
  if (!ctx || !out || n_edges < 0 || (n_edges > 0 && !edges)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::SdbgView g;
    mf::dev_seq2sdbg(ctx->c, edges, n_edges, mf::SeqsView{}, k, tip_mode, &g);
    ctx->c.end_call();
    fill(out, g, ctx->c);
  });
}
int mfsdbg_dev_read2sdbg(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t min_count, mfsdbg_dev_sdbg *out) {
  if (!ctx || !out) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::EdgesView e;
    mf::dev_count(ctx->c, view(reads), k, min_count, &e, nullptr);
    mf::SdbgView g;
    mf::dev_seq2sdbg(ctx->c, e.edges, e.n_edges, mf::SeqsView{}, k, 1, &g);
    ctx->c.end_call();
    fill(out, g, ctx->c);
  });
}
int mfsdbg_dev_pack_fastq(mfsdbg_ctx *ctx, const uint8_t *text, int64_t n_bytes, int32_t n_policy, mfsdbg_dev_reads *out,
                          int32_t *max_len) {
  if (!ctx || !out || n_bytes < 0 || (n_bytes > 0 && !text)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::ReadsView r;
    int ml = 0;
    mf::dev_pack_fastq(ctx->c, &text, &n_bytes, 1, n_policy, &r, &ml);
    ctx->c.end_call();
    out->packed = r.packed;
    out->starts = r.starts;
    out->n_reads = r.n_reads;
    out->n_bases = r.n_bases;
    if (max_len) *max_len = ml;
  });
}
int mfsdbg_dev_count_hist(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t l1_bits, uint64_t *hist_dev) {
  if (!ctx || !hist_dev) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_count_hist(ctx->c, view(reads), k, l1_bits, reinterpret_cast<unsigned long long *>(hist_dev));
    ctx->c.end_call();
  });
}
int mfsdbg_dev_count_scatter(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t l1_bits, const uint64_t *hist_dev,
                             uint32_t *keys_out, int64_t capacity) {
  if (!ctx || !hist_dev || capacity < 0 || (!keys_out && capacity > 0)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_count_scatter(ctx->c, view(reads), k, l1_bits, reinterpret_cast<const unsigned long long *>(hist_dev), keys_out,
                          capacity, nullptr);
    ctx->c.end_call();
  });
}
int mfsdbg_dev_count_finish(mfsdbg_ctx *ctx, uint32_t *keys, uint32_t *scratch, int64_t n_keys, const int64_t *chunk_start,
                            const int64_t *chunk_size, const int32_t *chunk_seg, int32_t n_chunks, int32_t n_segs, int32_t k,
                            int32_t l1_bits, int32_t min_count, mfsdbg_dev_edges *out, int64_t *counting_host) {
  if (!ctx || !out || n_keys < 0 || n_chunks < 0 || n_segs < 1) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::EdgesView e;
    mf::dev_count_finish(ctx->c, keys, scratch, n_keys, chunk_start, chunk_size, chunk_seg, n_chunks, n_segs, k, l1_bits,
                         min_count, &e, counting_host);
    ctx->c.end_call();
    fill(out, e);
  });
}
int mfsdbg_dev_sdbg_items(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int32_t k, uint32_t *items_out) {
  if (!ctx || n_edges < 0 || (n_edges > 0 && (!edges || !items_out))) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_sdbg_items(ctx->c, edges, n_edges, k, items_out);
    ctx->c.end_call();
  });
}
int mfsdbg_dev_records_hist(mfsdbg_ctx *ctx, const uint32_t *records, int64_t n, int32_t words, int32_t l1_bits, uint64_t *hist_dev) {
  if (!ctx || !hist_dev || n < 0 || (n > 0 && !records)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_records_hist(ctx->c, records, n, words, l1_bits, reinterpret_cast<unsigned long long *>(hist_dev));
    ctx->c.end_call();
  });
}
int mfsdbg_dev_records_scatter(mfsdbg_ctx *ctx, const uint32_t *records, int64_t n, int32_t words, int32_t l1_bits,
                               const uint64_t *hist_dev, uint32_t *out) {
  if (!ctx || !hist_dev || n < 0 || (n > 0 && (!records || !out))) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_records_scatter(ctx->c, records, n, words, l1_bits, reinterpret_cast<const unsigned long long *>(hist_dev), out, nullptr);
    ctx->c.end_call();
  });
}
int mfsdbg_dev_sdbg_finish(mfsdbg_ctx *ctx, uint32_t *items, uint32_t *scratch, int64_t n_items, const int64_t *chunk_start,
                           const int64_t *chunk_size, const int32_t *chunk_seg, int32_t n_chunks, int32_t n_segs, int32_t k,
                           int32_t l1_bits, int32_t tip_mode, mfsdbg_dev_sdbg *out) {
  if (!ctx || !out || n_items < 0 || n_chunks < 0 || n_segs < 1) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::SdbgView g;
    mf::dev_sdbg_finish(ctx->c, items, scratch, n_items, chunk_start, chunk_size, chunk_seg, n_chunks, n_segs, k, l1_bits, tip_mode,
                        &g);
    ctx->c.end_call();
    fill(out, g, ctx->c);
  });
}
int32_t mfsdbg_words_per_item(int32_t k) { return mf::words_item(k); }

// ---- fused partition + exchange over peer memory (NVLink): buffers the library allocates are exportable through CUDA IPC
int mfsdbg_dev_alloc(mfsdbg_ctx *ctx, uint64_t bytes, void **out) {
  if (!ctx || !out) return MFSDBG_EINVAL;
  return guarded([&] {
    MF_CUDA(cudaSetDevice(ctx->c.device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 256);
    if (e != cudaSuccess) {
      cudaGetLastError();
      throw mf::CudaError("out of device memory allocating " + std::to_string(bytes >> 20) + " MiB");
    }
  });
}
int mfsdbg_dev_free(mfsdbg_ctx *ctx, void *ptr) {
  if (!ctx) return MFSDBG_EINVAL;
  return guarded([&] {
    MF_CUDA(cudaSetDevice(ctx->c.device));
    MF_CUDA(cudaStreamSynchronize(ctx->c.stream));
    if (ptr) MF_CUDA(cudaFree(ptr));
  });
}
int mfsdbg_ipc_export(mfsdbg_ctx *ctx, void *ptr, uint8_t *handle64) {
  if (!ctx || !ptr || !handle64) return MFSDBG_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  return guarded([&] {
    MF_CUDA(cudaSetDevice(ctx->c.device));
    cudaIpcMemHandle_t h;
    MF_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
  });
}
int mfsdbg_ipc_open(mfsdbg_ctx *ctx, const uint8_t *handle64, void **out) {
  if (!ctx || !handle64 || !out) return MFSDBG_EINVAL;
  return guarded([&] {
    MF_CUDA(cudaSetDevice(ctx->c.device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    MF_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  });
}
int mfsdbg_ipc_close(mfsdbg_ctx *ctx, void *ptr) {
  if (!ctx) return MFSDBG_EINVAL;
  return guarded([&] {
    MF_CUDA(cudaSetDevice(ctx->c.device));
    if (ptr) MF_CUDA(cudaIpcCloseMemHandle(ptr));
  });
}
int mfsdbg_dev_count_scatter_peer(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t l1_bits, const uint64_t *bin_base_dev) {
  if (!ctx || !bin_base_dev) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_count_scatter(ctx->c, view(reads), k, l1_bits, nullptr, nullptr, -1, reinterpret_cast<const unsigned long long *>(bin_base_dev));
    ctx->c.end_call();
  });
}
int mfsdbg_dev_records_scatter_peer(mfsdbg_ctx *ctx, const uint32_t *records, int64_t n, int32_t words, int32_t l1_bits,
                                    const uint64_t *bin_base_dev) {
  if (!ctx || !bin_base_dev || n < 0 || (n > 0 && !records)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_records_scatter(ctx->c, records, n, words, l1_bits, nullptr, nullptr, reinterpret_cast<const unsigned long long *>(bin_base_dev));
    ctx->c.end_call();
  });
}
// ---- super-k-mer exchange (skm.cu)
int32_t mfsdbg_skm_supported(int32_t k) { return mf::skm_supported(k) ? 1 : 0; }
int64_t mfsdbg_skm_key_capacity(int64_t n_keys) { return mf::skm_key_capacity(n_keys < 0 ? 0 : n_keys); }
int mfsdbg_dev_skm_scatter(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t n_dst, const uint64_t *dst_ptrs,
                           const int64_t *dst_caps, int64_t stride, int64_t *counts_out) {
  if (!ctx || !reads || !counts_out || (dst_ptrs && !dst_caps)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_skm_scatter(ctx->c, view(reads), k, n_dst, dst_ptrs, dst_caps, stride, counts_out);
    ctx->c.end_call();
  });
}
int mfsdbg_dev_count_skm(mfsdbg_ctx *ctx, const uint64_t *records, const int64_t *chunk_start, const int64_t *chunk_size,
                         int32_t n_chunks, int64_t n_keys, int32_t k, int32_t min_count, uint32_t *keys, uint32_t *scratch,
                         int64_t capacity, mfsdbg_dev_edges *out) {
  if (!ctx || !out || n_chunks < 0 || n_keys < 0 || capacity < 0 || (n_chunks > 0 && (!chunk_start || !chunk_size))) return MFSDBG_EINVAL;
  if (n_keys > 0 && (!records || !keys || !scratch)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::EdgesView e;
    mf::dev_count_skm(ctx->c, records, chunk_start, chunk_size, n_chunks, n_keys, k, min_count, keys, scratch, capacity, &e);
    ctx->c.end_call();
    fill(out, e);
  });
}
// ---- the item filter across GPUs (ksdist.cu)
int32_t mfsdbg_ks_supported(int32_t k) { return mf::ksd_supported(k) ? 1 : 0; }
int mfsdbg_ks_geometry(int64_t n_edges_global, int32_t world, int32_t *log_slots, int32_t *slice_log) {
  if (!log_slots || !slice_log || n_edges_global < 0 || world < 1) return MFSDBG_EINVAL;
  int a = 0, b = 0;
  mf::ksd_geometry(n_edges_global, world, &a, &b);
  *log_slots = a;
  *slice_log = b;
  return MFSDBG_OK;
}
int mfsdbg_dev_ks_hist(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int32_t k, int32_t log_slots, int32_t slice_log,
                       uint64_t *hist_dev) {
  if (!ctx || !hist_dev || n_edges < 0 || (n_edges > 0 && !edges)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_ksd_hist(ctx->c, edges, n_edges, k, log_slots, slice_log, reinterpret_cast<unsigned long long *>(hist_dev));
    ctx->c.end_call();
  });
}
int mfsdbg_dev_ks_scatter_peer(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int32_t k, int32_t log_slots, int32_t slice_log,
                               const uint64_t *bin_base_dev) {
  if (!ctx || !bin_base_dev || n_edges < 0 || (n_edges > 0 && !edges)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::dev_ksd_scatter(ctx->c, edges, n_edges, k, log_slots, slice_log, reinterpret_cast<const unsigned long long *>(bin_base_dev));
    ctx->c.end_call();
  });
}
int mfsdbg_dev_ks_filter(mfsdbg_ctx *ctx, const uint64_t *inserts, int64_t n_inserts, const uint64_t *queries, int64_t n_queries,
                         int32_t log_slots, int32_t slice_log, int32_t slice_lo, int32_t n_owned, int64_t *n_miss) {
  if (!ctx || !n_miss || n_inserts < 0 || n_queries < 0 || (n_inserts > 0 && !inserts) || (n_queries > 0 && !queries)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    *n_miss = mf::dev_ksd_filter(ctx->c, inserts, n_inserts, queries, n_queries, log_slots, slice_log, slice_lo, n_owned);
    ctx->c.end_call();
  });
}
int mfsdbg_dev_ks_items(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int64_t n_miss, int32_t k, uint32_t *items_out,
                        int64_t capacity, int64_t *n_items) {
  if (!ctx || !n_items || n_edges < 0 || n_miss < 0 || capacity < 0 || (n_edges > 0 && !edges) || (capacity > 0 && !items_out)) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    *n_items = mf::dev_ksd_items(ctx->c, edges, n_edges, n_miss, k, items_out, capacity);
    ctx->c.end_call();
  });
}
int mfsdbg_dev_synth_reads(mfsdbg_ctx *ctx, const mfsdbg_synth_spec *spec, mfsdbg_dev_reads *out) {
  if (!ctx || !spec || !out) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    ctx->c.begin_call();
    mf::ReadsView r;
    mf::dev_synth(ctx->c, *spec, &r);
    ctx->c.end_call();
    out->packed = r.packed;
    out->starts = r.starts;
    out->n_reads = r.n_reads;
    out->n_bases = r.n_bases;
  });
}

int mfsdbg_host_read2sdbg(mfsdbg_ctx *ctx, const uint32_t *packed_host, const int64_t *starts_host, int64_t n_reads,
                          int64_t n_bases, int32_t k, int32_t min_count, mfsdbg_host_sdbg *out) {
  if (!ctx || !out || n_reads < 0 || n_bases < 0 || (n_bases > 0 && (!packed_host || !starts_host))) return MFSDBG_EINVAL;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    mf::Ctx &c = ctx->c;
    c.begin_call();
    const size_t wbytes = (size_t)((n_bases + 15) >> 4) * 4, sbytes = sizeof(int64_t) * (size_t)(n_reads + 1);
    mf::EdgesView e;
    // large inputs: chunked transfer, the reads-fed partition level follows the chunks as they land
    const char *minb = getenv("MFSDBG_H2D_MIN_BASES");
    if (!(n_bases >= (minb ? atoll(minb) : (long long)(64 << 20)) && mf::dev_count_host(c, packed_host, starts_host, n_reads, n_bases, k, min_count, &e))) {
      c.in_words.reserve(wbytes + 64);
      c.in_starts.reserve(sbytes);
      {
        mf::Stage st(c, "h2d");
        MF_CUDA(cudaMemsetAsync((char *)c.in_words.p + wbytes, 0, 64, c.stream));
        if (wbytes) MF_CUDA(cudaMemcpyAsync(c.in_words.p, packed_host, wbytes, cudaMemcpyHostToDevice, c.stream));
        MF_CUDA(cudaMemcpyAsync(c.in_starts.p, starts_host, sbytes, cudaMemcpyHostToDevice, c.stream));
      }
      mf::ReadsView r{c.in_words.as<uint32_t>(), c.in_starts.as<int64_t>(), n_reads, n_bases};
      mf::dev_count(c, r, k, min_count, &e, nullptr);
    }
    mf::SdbgView g;
    mf::dev_seq2sdbg(c, e.edges, e.n_edges, mf::SeqsView{}, k, 1, &g);
    // over PCIe the records travel as megahit's 16-bit packed items plus the few multiplicities beyond 254 (half the bytes)
    const size_t rb = (size_t)g.n_items * 2, lb = (size_t)g.n_tips * g.words_tip * 4, pb = (size_t)g.n_large * 8;
    c.slab_reserve(rb + pb + 4096);
    c.slab_reset();
    uint16_t *d_rec16 = c.alloc<uint16_t>((size_t)g.n_items + 8);
    unsigned long long *d_pairs = c.alloc<unsigned long long>((size_t)g.n_large + 1), *d_cur = c.alloc<unsigned long long>(1);
    c.out_rec.reserve(rb + 64);
    c.out_labels.reserve(lb + 64);
    c.out_large.reserve(pb + 64);
    {
      mf::Stage st(c, "d2h");
      mf::dev_sdbg_pack16(c, g, d_rec16, d_pairs, d_cur);
      if (rb) MF_CUDA(cudaMemcpyAsync(c.out_rec.p, d_rec16, rb, cudaMemcpyDeviceToHost, c.stream));
      if (lb) MF_CUDA(cudaMemcpyAsync(c.out_labels.p, g.labels, lb, cudaMemcpyDeviceToHost, c.stream));
      if (pb) MF_CUDA(cudaMemcpyAsync(c.out_large.p, d_pairs, pb, cudaMemcpyDeviceToHost, c.stream));
    }
    c.end_call();
    {
      unsigned long long *pairs = c.out_large.as<unsigned long long>();
      std::sort(pairs, pairs + g.n_large);
      c.out_large_index.resize((size_t)g.n_large);
      c.out_large_mult.resize((size_t)g.n_large);
      for (int64_t i = 0; i < g.n_large; ++i) {
        c.out_large_index[(size_t)i] = (int64_t)(pairs[i] >> 16);
        c.out_large_mult[(size_t)i] = (uint16_t)(pairs[i] & 0xffffu);
      }
    }
    out->rec = c.out_rec.as<uint16_t>();
    out->tip_labels = c.out_labels.as<uint32_t>();
    out->large_index = c.out_large_index.data();
    out->large_mult = c.out_large_mult.data();
    out->n_items = g.n_items;
    out->n_tips = g.n_tips;
    out->n_large = g.n_large;
    out->k = g.k;
    out->words_per_tip = g.words_tip;
    out->h2d_bytes = (int64_t)(wbytes + sbytes);
    out->d2h_bytes = (int64_t)(rb + lb + pb);
  });
}

int mfsdbg_dev_copy(mfsdbg_ctx *ctx, void *dst, const void *src, uint64_t bytes, int32_t kind) {
  if (!ctx || kind < 0 || kind > 2 || (bytes && (!dst || !src))) return MFSDBG_EINVAL;
  return guarded([&] {
    MF_CUDA(cudaSetDevice(ctx->c.device));
    const cudaMemcpyKind kk = kind == 0 ? cudaMemcpyDeviceToHost : kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (bytes) MF_CUDA(cudaMemcpyAsync(dst, src, bytes, kk, ctx->c.stream));
    MF_CUDA(cudaStreamSynchronize(ctx->c.stream));
  });
}
int mfsdbg_ctx_edge_bucket_counts(mfsdbg_ctx *ctx, int64_t *out) {
  if (!ctx || !out) return MFSDBG_EINVAL;
  memcpy(out, ctx->c.edge_bucket_counts.data(), sizeof(int64_t) * mf::kNumBuckets);
  return MFSDBG_OK;
}
int mfsdbg_ctx_sdbg_bucket_stats(mfsdbg_ctx *ctx, int64_t *out) {
  if (!ctx || !out) return MFSDBG_EINVAL;
  memcpy(out, ctx->c.sdbg_bucket_stats.data(), sizeof(int64_t) * mf::kNumBuckets * 3);
  return MFSDBG_OK;
}

// ---- file-level entry points ------------------------------------------------------------------
static int pick_device(const mfsdbg_opts *o) {
  if (o && o->n_gpus > 0 && o->gpu_ids) return o->gpu_ids[0];
  return 0;
}
// n_gpus > 1: the listed devices (or 0 .. n_gpus-1) share the job inside this one process (multi.cu)
static std::vector<int> device_list(const mfsdbg_opts *o) {
  std::vector<int> d;
  for (int i = 0; i < o->n_gpus; ++i) d.push_back(o->gpu_ids ? o->gpu_ids[i] : i);
  return d;
}
static int need_device() {
  if (mfsdbg_device_count() < 1) {
    g_err = "no CUDA device visible: libmfsdbg has no CPU fallback";
    return MFSDBG_ENODEV;
  }
  return MFSDBG_OK;
}
int mfsdbg_buildlib(const char *lib_file, const char *out_prefix, int32_t n_policy) {
  if (!lib_file || !out_prefix) return MFSDBG_EINVAL;
  if (int rc = need_device()) return rc;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    mf::Ctx c(0);
    mf::file_buildlib(c, lib_file, out_prefix, n_policy);
  });
}
int mfsdbg_count(const mfsdbg_opts *o) {
  if (!o || !o->read_lib_file || !o->output_prefix) { g_err = "count needs --read_lib_file and --output_prefix"; return MFSDBG_EINVAL; }
  if (o->need_mercy) { g_err = "--need_mercy is not supported (MitoFlex never passes it: configurations.py:67)"; return MFSDBG_EINVAL; }
  if (int rc = need_device()) return rc;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    if (o->n_gpus > 1) return mf::file_count_multi(device_list(o), o->read_lib_file, o->k, o->min_count, o->output_prefix);
    mf::Ctx c(pick_device(o));
    mf::file_count(c, o->read_lib_file, o->k, o->min_count, o->output_prefix, std::max(1, o->num_cpu_threads));
  });
}
int mfsdbg_seq2sdbg(const mfsdbg_opts *o) {
  if (!o || !o->output_prefix) { g_err = "seq2sdbg needs --output_prefix"; return MFSDBG_EINVAL; }
  if (o->need_mercy) { g_err = "--need_mercy is not supported (MitoFlex never passes it: configurations.py:67)"; return MFSDBG_EINVAL; }
  if (int rc = need_device()) return rc;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    if (o->n_gpus > 1)
      return mf::file_seq2sdbg_multi(device_list(o), o->k, o->kmer_from, o->input_prefix, o->contig, o->bubble, o->addi_contig,
                                     o->local_contig, o->output_prefix);
    mf::Ctx c(pick_device(o));
    mf::file_seq2sdbg(c, o->k, o->kmer_from, o->input_prefix, o->contig, o->bubble, o->addi_contig, o->local_contig,
                      o->output_prefix, std::max(1, o->num_cpu_threads));
  });
}
int mfsdbg_read2sdbg(const mfsdbg_opts *o) {
  if (!o || !o->read_lib_file || !o->output_prefix) { g_err = "read2sdbg needs --read_lib_file and --output_prefix"; return MFSDBG_EINVAL; }
  if (o->need_mercy) { g_err = "--need_mercy is not supported (MitoFlex never passes it: configurations.py:67)"; return MFSDBG_EINVAL; }
  if (int rc = need_device()) return rc;
  std::lock_guard<std::mutex> lk(g_job_mutex);
  return guarded([&] {
    if (o->n_gpus > 1) return mf::file_read2sdbg_multi(device_list(o), o->read_lib_file, o->k, o->min_count, o->output_prefix);
    mf::Ctx c(pick_device(o));
    mf::file_read2sdbg(c, o->read_lib_file, o->k, o->min_count, o->output_prefix, std::max(1, o->num_cpu_threads));
  });
}
