// hostio.h -- file-level side of libmfsdbg: the on-disk contract of megahit_core's sub-commands
// (reads.lib / .bin / .lib_info, <prefix>.edges.*, <prefix>.sdbg.*, contig FASTA).
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "engine.cuh"

namespace mf {

struct IoError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

void file_buildlib(Ctx &c, const char *lib_file, const char *out_prefix, int n_policy);
void file_count(Ctx &c, const char *read_lib_file, int k, int min_count, const char *out_prefix, int n_files);
void file_seq2sdbg(Ctx &c, int k, int k_from, const char *input_prefix, const char *contig, const char *bubble,
                   const char *addi_contig, const char *local_contig, const char *out_prefix, int n_files);
void file_read2sdbg(Ctx &c, const char *read_lib_file, int k, int min_count, const char *out_prefix, int n_files);

// multi.cu -- the same sub-commands on several GPUs of one box in ONE process (one host thread + context per device, peer
// access): rank r ends with a contiguous piece of the globally sorted streams
struct MultiRank;
struct MultiGpu {
  std::vector<std::unique_ptr<MultiRank>> ranks;
  explicit MultiGpu(const std::vector<int> &devices);
  ~MultiGpu();
  int world() const;
  Ctx &ctx(int r);
  const EdgesView &edges(int r) const;
  const SdbgView &sdbg(int r) const;
  const std::vector<int64_t> &counting(int r) const;
  void count(const std::vector<ReadsView> &reads, int k, int min_count, bool want_counting);
  void read2sdbg(const std::vector<ReadsView> &reads, int k, int min_count);
  void seq2sdbg(const std::vector<const uint32_t *> &edges, const std::vector<int64_t> &n_edges, const std::vector<SeqsView> &seqs, int k);
};
void file_count_multi(const std::vector<int> &devices, const char *read_lib_file, int k, int min_count, const char *out_prefix);
void file_read2sdbg_multi(const std::vector<int> &devices, const char *read_lib_file, int k, int min_count, const char *out_prefix);
void file_seq2sdbg_multi(const std::vector<int> &devices, int k, int k_from, const char *input_prefix, const char *contig, const char *bubble,
                         const char *addi_contig, const char *local_contig, const char *out_prefix);

// pack.cu
void dev_pack_fastq(Ctx &c, const uint8_t *const *texts, const int64_t *n_bytes, int n_texts, int n_policy, ReadsView *out,
                    int *max_len);
void reads_to_bin_stream(Ctx &c, const ReadsView &r, std::vector<uint32_t> *stream);
void bin_stream_to_reads(Ctx &c, const uint32_t *stream, int64_t n_words_stream, ReadsView *out);

}  // namespace mf
