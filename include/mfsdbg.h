/*
 * mfsdbg.h -- C ABI of libmfsdbg.so: B200-native (sm_100a) succinct de Bruijn graph construction,
 * a drop-in for the `megahit_core buildlib | count | seq2sdbg | read2sdbg` sub-commands that MitoFlex
 * runs at every k of its k-list.
 *
 * Reference interfaces replaced (all in /root/reference, megahit itself is an un-vendored conda
 * dependency pinned by environment.yml:8):
 *   mfsdbg_buildlib   <- shell_call(MEGAHIT_CORE, 'buildlib', read_lib, read_lib)   assemble/assemble_wrapper.py:193
 *   mfsdbg_count      <- shell_call(MEGAHIT_CORE, 'count', **count_opts)            assemble/assemble_wrapper.py:215-224
 *   mfsdbg_seq2sdbg   <- shell_call(MEGAHIT_CORE, 'seq2sdbg', **options)            assemble/assemble_wrapper.py:226-258
 *   mfsdbg_read2sdbg  <- megahit_core read2sdbg (the 1-pass route `one_pass` names, configurations.py:72;
 *                        assemble_wrapper.py:216 only skips `count`, the sub-command itself is megahit's)
 * The option struct mirrors the argv that utility/helper.py:50-75 (concat_command) builds from the
 * option dicts of assemble_wrapper.py:204-250.
 *
 * Conventions: plain C, caller owns every pointer it passes for the duration of the call, the library
 * owns everything it allocates.  Every function returns 0 on success or a negative MFSDBG_E* code;
 * mfsdbg_last_error() returns a thread-local message.  No exceptions and no exit() cross this boundary.
 * There is no CPU fallback: without a CUDA device every compute entry point returns MFSDBG_ENODEV.
 */
#ifndef MFSDBG_H
#define MFSDBG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MFSDBG_VERSION 100

#define MFSDBG_OK 0
#define MFSDBG_EINVAL (-1)   /* bad argument / unsupported option          */
#define MFSDBG_EIO (-2)      /* file missing, unreadable or malformed      */
#define MFSDBG_ENODEV (-3)   /* no usable CUDA device                      */
#define MFSDBG_ECUDA (-4)    /* CUDA runtime / kernel failure              */
#define MFSDBG_ENOMEM (-5)   /* host or device memory exhausted            */
#define MFSDBG_EINTERNAL (-6)

/* N policy of the read packer (mfsdbg_buildlib / mfsdbg_dev_pack_fastq) */
#define MFSDBG_N_MEGAHIT 0   /* megahit FastxReader::TrimN: keep the first N-free segment of a read */
#define MFSDBG_N_SPLIT 1     /* every N-free segment becomes its own read                            */

/* ---- file-level entry points: what the megahit_core sub-commands do ------------------------- */
typedef struct mfsdbg_opts {
  int32_t k;                 /* -k                                                        */
  int32_t kmer_from;         /* --kmer_from                                               */
  int32_t min_count;         /* -m (count / read2sdbg)                                    */
  int32_t mem_flag;          /* --mem_flag      (accepted, sizing is by HBM instead)      */
  int32_t num_cpu_threads;   /* --num_cpu_threads (host I/O threads; also output files)   */
  int32_t need_mercy;        /* --need_mercy    (unsupported -> MFSDBG_EINVAL if nonzero) */
  int64_t host_mem;          /* --host_mem      (accepted)                                */
  int32_t n_gpus;            /* 0 = use device 0                                          */
  const int32_t *gpu_ids;    /* n_gpus entries or NULL                                    */
  int32_t n_policy;          /* buildlib only: MFSDBG_N_*                                 */
  const char *read_lib_file; /* --read_lib_file                                           */
  const char *input_prefix;  /* --input_prefix                                            */
  const char *output_prefix; /* --output_prefix                                           */
  const char *contig;        /* --contig                                                  */
  const char *bubble;        /* --bubble                                                  */
  const char *addi_contig;   /* --addi_contig                                             */
  const char *local_contig;  /* --local_contig                                            */
} mfsdbg_opts;

int mfsdbg_version(void);
int mfsdbg_device_count(void);
const char *mfsdbg_last_error(void);

/* megahit_core buildlib <lib_file> <out_prefix>: writes <out_prefix>.bin and <out_prefix>.lib_info */
int mfsdbg_buildlib(const char *lib_file, const char *out_prefix, int32_t n_policy);
/* megahit_core count: <output_prefix>.edges.<i>, .edges.info, .counting */
int mfsdbg_count(const mfsdbg_opts *opts);
/* megahit_core seq2sdbg: <output_prefix>.sdbg.<i>, .sdbg_info */
int mfsdbg_seq2sdbg(const mfsdbg_opts *opts);
/* megahit_core read2sdbg: same outputs as seq2sdbg, straight from the read library */
int mfsdbg_read2sdbg(const mfsdbg_opts *opts);

/* ---- device-level entry points: buffers already in HBM (bench, multi-GPU driver, tests) ------ */
typedef struct mfsdbg_ctx mfsdbg_ctx;

/* one context per process and GPU; owns a stream and a workspace slab */
mfsdbg_ctx *mfsdbg_ctx_create(int32_t device);
void mfsdbg_ctx_destroy(mfsdbg_ctx *ctx);
/* cap the workspace (bytes); 0 = 85 % of the device memory free at first use */
int mfsdbg_ctx_set_mem_limit(mfsdbg_ctx *ctx, uint64_t bytes);
/* the stream every kernel of this context is launched on (a cudaStream_t) */
void *mfsdbg_ctx_stream(mfsdbg_ctx *ctx);
/* run on a caller-owned stream instead (e.g. the host framework's current stream, so that the caller's CUDA
 * events bracket the library's kernels); NULL restores the context's own stream */
int mfsdbg_ctx_set_stream(mfsdbg_ctx *ctx, void *cuda_stream);
/* kernels launched by this context since creation (for bench.py's gpu_launches) */
int64_t mfsdbg_ctx_launches(mfsdbg_ctx *ctx);
/* per-stage device milliseconds of the last call, as "name=ms;name=ms;..." (profiling must be enabled) */
int mfsdbg_ctx_set_profiling(mfsdbg_ctx *ctx, int32_t on);
const char *mfsdbg_ctx_last_profile(mfsdbg_ctx *ctx);

/* Packed reads in HBM: 2 bits per base (A=0 C=1 G=2 T=3), 16 bases per uint32, first base in the top bits,
 * reads back to back; starts[n_reads+1] are base offsets.  The packed array must be padded with 64 readable
 * bytes past ceil(n_bases/16) words. */
typedef struct mfsdbg_dev_reads {
  const uint32_t *packed;   /* device */
  const int64_t *starts;    /* device, n_reads + 1 */
  int64_t n_reads;
  int64_t n_bases;
} mfsdbg_dev_reads;

/* Sorted solid edges in HBM (the logical content of <prefix>.edges.*): n_edges records of words_per_edge
 * uint32, multiplicity in the low 16 bits of the last word.  Owned by the context until the next
 * mfsdbg_dev_* call that produces edges, or mfsdbg_ctx_destroy. */
typedef struct mfsdbg_dev_edges {
  const uint32_t *edges;    /* device */
  int64_t n_edges;
  int32_t k;
  int32_t words_per_edge;
  int64_t n_keys;           /* (k+1)-mer occurrences counted */
} mfsdbg_dev_edges;

/* The sdbg in HBM (logical content of <prefix>.sdbg.*), items in BOSS order:
 * rec[i] = w | last<<4 | tip<<5 | multiplicity<<8 ; tip labels (words_per_tip uint32 each) in item order. */
typedef struct mfsdbg_dev_sdbg {
  const uint32_t *rec;         /* device, n_items */
  const uint32_t *tip_labels;  /* device, n_tips * words_per_tip */
  const int64_t *bucket_items; /* device, 65536: items per megahit bucket (first 8 bases) */
  int64_t n_items, n_tips, n_large;
  int32_t k, words_per_tip;
} mfsdbg_dev_sdbg;

/* count: reads -> sorted solid (k+1)-mer edges.  counting_host (65536 int64, may be NULL) receives the
 * distinct-edge multiplicity histogram that megahit writes to <prefix>.counting. */
int mfsdbg_dev_count(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t min_count,
                     mfsdbg_dev_edges *out, int64_t *counting_host);
/* seq2sdbg on device-resident edges (k_min iteration; contigs go through mfsdbg_seq2sdbg). */
int mfsdbg_dev_seq2sdbg(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int32_t k, int32_t tip_mode,
                        mfsdbg_dev_sdbg *out);
/* read2sdbg = count + seq2sdbg fused in HBM (no edge files), tip labels in megahit's stage-2 layout. */
int mfsdbg_dev_read2sdbg(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t min_count,
                         mfsdbg_dev_sdbg *out);

/* FASTQ text in HBM -> packed reads (K1).  text = the raw file bytes (4-line records); the library
 * allocates the outputs inside the context (valid until the next pack call). */
int mfsdbg_dev_pack_fastq(mfsdbg_ctx *ctx, const uint8_t *text, int64_t n_bytes, int32_t n_policy,
                          mfsdbg_dev_reads *out, int32_t *max_len);

/* ---- staged count for the multi-GPU driver (reads sharded by GPU, keys routed by prefix) ------ */
/* l1_bits-bit prefix histogram of this GPU's keys: hist_dev[1 << l1_bits] uint64 (device, zeroed by callee) */
int mfsdbg_dev_count_hist(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t l1_bits,
                          uint64_t *hist_dev);
/* keys of this GPU's reads, partitioned by l1_bits-bit prefix into keys_out (device, capacity in records);
 * bin b occupies [bin_start[b], bin_start[b+1]) where bin_start is the exclusive prefix of this GPU's hist. */
int mfsdbg_dev_count_scatter(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t l1_bits,
                             const uint64_t *hist_dev, uint32_t *keys_out, int64_t capacity);
/* finish: `keys` holds n_chunks chunks (chunk_start/size in records, chunk_seg = prefix bin - seg_base) that all
 * share their l1_bits-bit prefix per segment; sorts, counts, filters.  scratch must hold as many records as keys. */
int mfsdbg_dev_count_finish(mfsdbg_ctx *ctx, uint32_t *keys, uint32_t *scratch, int64_t n_keys,
                            const int64_t *chunk_start, const int64_t *chunk_size, const int32_t *chunk_seg,
                            int32_t n_chunks, int32_t n_segs, int32_t k, int32_t l1_bits, int32_t min_count,
                            mfsdbg_dev_edges *out, int64_t *counting_host);
/* staged seq2sdbg, same shape: items of this GPU's edges -> (the driver exchanges them by prefix) -> finish.
 * items_out holds 6 * n_edges records of mfsdbg_words_per_item(k) words. */
int mfsdbg_dev_sdbg_items(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int32_t k, uint32_t *items_out);
/* histogram / partition of fixed-width records by their top l1_bits (bin b lands at the exclusive prefix of hist) */
int mfsdbg_dev_records_hist(mfsdbg_ctx *ctx, const uint32_t *records, int64_t n, int32_t words, int32_t l1_bits,
                            uint64_t *hist_dev);
int mfsdbg_dev_records_scatter(mfsdbg_ctx *ctx, const uint32_t *records, int64_t n, int32_t words, int32_t l1_bits,
                               const uint64_t *hist_dev, uint32_t *out);
int mfsdbg_dev_sdbg_finish(mfsdbg_ctx *ctx, uint32_t *items, uint32_t *scratch, int64_t n_items,
                           const int64_t *chunk_start, const int64_t *chunk_size, const int32_t *chunk_seg,
                           int32_t n_chunks, int32_t n_segs, int32_t k, int32_t l1_bits, int32_t tip_mode,
                           mfsdbg_dev_sdbg *out);
int32_t mfsdbg_words_per_item(int32_t k);

/* ---- fused partition + exchange over NVLink peer memory ---------------------------------------------------
 * The owner GPU allocates its receive buffer with mfsdbg_dev_alloc, exports it (CUDA IPC, 64-byte handle), the
 * other ranks open it; the *_scatter_peer kernels then write every record of prefix bin b at byte address
 * bin_base_dev[b] + (running count of bin b) * record size -- directly into the owner's HBM, no staging copy and no
 * separate all-to-all.  bin_base_dev is a device array of 1 << l1_bits addresses. */
int mfsdbg_dev_alloc(mfsdbg_ctx *ctx, uint64_t bytes, void **out);
int mfsdbg_dev_free(mfsdbg_ctx *ctx, void *ptr);
int mfsdbg_ipc_export(mfsdbg_ctx *ctx, void *ptr, uint8_t *handle64);
int mfsdbg_ipc_open(mfsdbg_ctx *ctx, const uint8_t *handle64, void **out);
int mfsdbg_ipc_close(mfsdbg_ctx *ctx, void *ptr);
int mfsdbg_dev_count_scatter_peer(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t l1_bits,
                                  const uint64_t *bin_base_dev);
int mfsdbg_dev_records_scatter_peer(mfsdbg_ctx *ctx, const uint32_t *records, int64_t n, int32_t words, int32_t l1_bits,
                                    const uint64_t *bin_base_dev);
int32_t mfsdbg_words_per_key(int32_t k);
int32_t mfsdbg_words_per_edge(int32_t k);

/* ---- super-k-mer exchange (16 <= k <= 26): what crosses NVLink in the multi-GPU count ----------------------
 * megahit counts in one address space (sorting/kmer_counter.cpp: every thread's lv1 scan sees every read); with the
 * reads sharded by GPU the (k+1)-mers have to reach the GPU that counts them.  Instead of one 8-byte key per
 * (k+1)-mer, a run of consecutive (k+1)-mers of a read whose minimizer hashes to the same destination travels as ONE
 * 64-bit record: the run's bases left-aligned (2 bits each, first base in the top bits, true orientation, at most 30)
 * and the run length - 1 in the low 3 bits.  A (k+1)-mer and its reverse complement share the minimizer, so every
 * canonical key is counted on exactly one GPU.
 *   skm_scatter : dst_ptrs[d] (host array of n_dst device addresses, peer mappings allowed) is where THIS source's records
 *                 for destination d go, dst_caps[d] how many fit; runs that would not fit are dropped and the returned
 *                 count exceeds the capacity (retry with more room).  dst_ptrs == NULL: count only, on every stride-th
 *                 tile of the reads (capacity estimate).  counts_out[0 .. n_dst): records, [n_dst .. 2 n_dst): (k+1)-mers.
 *   count_skm   : the single-GPU count over received records: `records` + chunks (start / size in records, one per
 *                 source), n_keys = the (k+1)-mers the sources announced; keys / scratch hold `capacity` records of 2 words
 *                 (mfsdbg_skm_key_capacity(n_keys)).  Result: this GPU's solid edges as mfsdbg_dev_count packs them, but in NO key
 *                 order (the count runs on bijectively mixed keys so that no minimizer crowds a bucket). */
int32_t mfsdbg_skm_supported(int32_t k);
int64_t mfsdbg_skm_key_capacity(int64_t n_keys);
int mfsdbg_dev_skm_scatter(mfsdbg_ctx *ctx, const mfsdbg_dev_reads *reads, int32_t k, int32_t n_dst, const uint64_t *dst_ptrs,
                           const int64_t *dst_caps, int64_t stride, int64_t *counts_out);
int mfsdbg_dev_count_skm(mfsdbg_ctx *ctx, const uint64_t *records, const int64_t *chunk_start, const int64_t *chunk_size,
                         int32_t n_chunks, int64_t n_keys, int32_t k, int32_t min_count, uint32_t *keys, uint32_t *scratch,
                         int64_t capacity, mfsdbg_dev_edges *out);

/* ---- the item filter of seq2sdbg across GPUs (16 <= k <= 31) ------------------------------------------------
 * megahit's SeqToSdbg writes 6 items per edge and lets Lv2Postprocess drop the "$" dummies that are shadowed by a real
 * edge; the single-GPU path asks a k-mer set instead and generates only the dummies that survive.  Across GPUs the hash
 * slices of that set are dealt out: slice s of 2^(log_slots - slice_log) belongs to rank s * world / nslices.
 *   ks_geometry     : table geometry for the GLOBAL edge count (all ranks must use the same)
 *   ks_hist         : this GPU's k-mer records per bin, bin = kind * nslices + slice (kind 0 = inserts, 1 = queries)
 *   ks_scatter_peer : record of bin b stored at bin_base_dev[b] + (running count of b) * 8 (peer addresses allowed)
 *   ks_filter       : the owner builds its slices [slice_lo, slice_lo + n_owned) from the received inserts and probes them with
 *                     the received queries (both slice-major); its share of the miss list stays in the context
 *   ks_items        : 2 real items per local edge + 2 dummies per miss of the last ks_filter -> items_out */
int32_t mfsdbg_ks_supported(int32_t k);
int mfsdbg_ks_geometry(int64_t n_edges_global, int32_t world, int32_t *log_slots, int32_t *slice_log);
int mfsdbg_dev_ks_hist(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int32_t k, int32_t log_slots, int32_t slice_log,
                       uint64_t *hist_dev);
int mfsdbg_dev_ks_scatter_peer(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int32_t k, int32_t log_slots,
                               int32_t slice_log, const uint64_t *bin_base_dev);
int mfsdbg_dev_ks_filter(mfsdbg_ctx *ctx, const uint64_t *inserts, int64_t n_inserts, const uint64_t *queries, int64_t n_queries,
                         int32_t log_slots, int32_t slice_log, int32_t slice_lo, int32_t n_owned, int64_t *n_miss);
int mfsdbg_dev_ks_items(mfsdbg_ctx *ctx, const uint32_t *edges, int64_t n_edges, int64_t n_miss, int32_t k, uint32_t *items_out,
                        int64_t capacity, int64_t *n_items);

/* ---- host-buffer entry point: what a caller holding the packed library in host memory uses ------------ */
/* The sdbg in (pinned, library-owned) host memory; valid until the next mfsdbg_host_* call on the context. */
typedef struct mfsdbg_host_sdbg {
  const uint16_t *rec;         /* host, n_items: w | last<<4 | tip<<5 | min(multiplicity, 255)<<8 -- megahit's packed sdbg item
                                  (SdbgWriter::Write); 255 = "large": the true multiplicity is in the side list below */
  const uint32_t *tip_labels;  /* host, n_tips * words_per_tip */
  const int64_t *large_index;  /* host, n_large: item indices with multiplicity > 254, ascending */
  const uint16_t *large_mult;  /* host, n_large: their multiplicities */
  int64_t n_items, n_tips, n_large;
  int32_t k, words_per_tip;
  int64_t h2d_bytes, d2h_bytes; /* bytes this call moved over PCIe */
} mfsdbg_host_sdbg;
/* read2sdbg with HOST inputs and outputs: copies packed reads (same layout as mfsdbg_dev_reads, host pointers;
 * pinned memory makes the copies asynchronous-fast) to HBM, builds the graph, copies the sdbg back. */
int mfsdbg_host_read2sdbg(mfsdbg_ctx *ctx, const uint32_t *packed_host, const int64_t *starts_host, int64_t n_reads,
                          int64_t n_bases, int32_t k, int32_t min_count, mfsdbg_host_sdbg *out);

/* plain copies on the context's stream, synchronous for the caller; kind: 0 = device->host, 1 = host->device,
 * 2 = device->device.  Lets ctypes hosts move results without a CUDA binding of their own. */
int mfsdbg_dev_copy(mfsdbg_ctx *ctx, void *dst, const void *src, uint64_t bytes, int32_t kind);
/* per-bucket tables of the last result held by the context (65536 entries each) */
int mfsdbg_ctx_edge_bucket_counts(mfsdbg_ctx *ctx, int64_t *out65536);
int mfsdbg_ctx_sdbg_bucket_stats(mfsdbg_ctx *ctx, int64_t *out65536x3);

/* ---- synthetic reads generated in HBM (bench / test tooling; SURVEY.md 8d generator) ---------- */
typedef struct mfsdbg_synth_spec {
  int64_t n_pairs;
  int32_t read_len;        /* 150 */
  int64_t mito_len;        /* 16500, circular */
  int64_t nuclear_len;     /* 50e6 */
  double mito_fraction;    /* 0.05 */
  double error_rate;       /* 0.005 */
  double n_rate;           /* 1e-4 per base; reads are cut at the first N (megahit TrimN) */
  double insert_mean, insert_sd;
  uint64_t seed;
} mfsdbg_synth_spec;
/* outputs allocated inside the context (valid until the next synth call) */
int mfsdbg_dev_synth_reads(mfsdbg_ctx *ctx, const mfsdbg_synth_spec *spec, mfsdbg_dev_reads *out);

#ifdef __cplusplus
}
#endif
#endif
