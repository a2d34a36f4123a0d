/*
 * oracle/mh_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the sDBG-construction path MitoFlex drives through
 * `megahit_core buildlib | count | seq2sdbg | read2sdbg`
 * (call sites: /root/reference/assemble/assemble_wrapper.py:193,224,258).
 *
 * PARITY UNPINNED: the algorithm lives in megahit v1.2.9 (bioconda pin in
 * /root/reference/environment.yml:8), which is NOT vendored under
 * /root/reference (only "assemble/megahit v1.2.9/LICENSE" is there), has no
 * binary in this image and cannot be fetched.  The reference also has no
 * tests or golden vectors for this path.  Everything here restates the
 * published megahit v1.2.9 algorithm (src/sorting/kmer_counter.cpp,
 * seq_to_sdbg.cpp, read_to_sdbg_s{1,2}.cpp, src/sequence/io/edge/,
 * src/sdbg/sdbg_{writer,meta}.h, src/sequence/io/) from recollection and is
 * pinned only by the hand-derived known-answer tests in tests/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef MH_ORACLE_H
#define MH_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUM_BUCKETS 65536
#define ORC_MAX_MUL 65535
#define ORC_SENTINEL 4

/* N policy of the read packer. 0 = megahit TrimN (keep the first N-free
 * segment), 1 = split every N-free segment into its own read. */
#define ORC_N_MEGAHIT 0
#define ORC_N_SPLIT 1

/* ---- reads ------------------------------------------------------------ */
typedef struct orc_reads orc_reads;
orc_reads *orc_reads_new(void);
void orc_reads_free(orc_reads *r);
/* append one read given as ASCII; applies the N policy. */
void orc_reads_add_ascii(orc_reads *r, const char *s, int64_t len, int n_policy);
int orc_reads_add_fastx(orc_reads *r, const char *path, int n_policy);
int orc_reads_add_fastx_pe(orc_reads *r, const char *p1, const char *p2, int n_policy);
orc_reads *orc_reads_load_bin(const char *bin_path);
int orc_reads_write_bin(const orc_reads *r, const char *bin_path);
int64_t orc_reads_count(const orc_reads *r);
int64_t orc_reads_bases(const orc_reads *r);
int orc_reads_max_len(const orc_reads *r);
const uint8_t *orc_reads_data(const orc_reads *r);   /* 1 byte/base, 0..3 */
const int64_t *orc_reads_starts(const orc_reads *r); /* count+1 offsets  */
/* megahit_core buildlib <lib_file> <out_prefix> */
int orc_cmd_buildlib(const char *lib_file, const char *out_prefix, int n_policy);

/* ---- count ------------------------------------------------------------ */
typedef struct orc_edges orc_edges;
orc_edges *orc_count(const uint8_t *bases, const int64_t *starts, int64_t n_reads,
                     int k, int min_count, int threads);
void orc_edges_free(orc_edges *e);
int orc_edges_k(const orc_edges *e);
int orc_edges_words(const orc_edges *e);
int orc_edges_sorted(const orc_edges *e);
int64_t orc_edges_n(const orc_edges *e);
const uint32_t *orc_edges_data(const orc_edges *e);
const int64_t *orc_edges_bucket_counts(const orc_edges *e); /* 65536 */
const int64_t *orc_edges_counting(const orc_edges *e);      /* 65536: [c] = #distinct edges seen c times (c capped) */
int orc_edges_write(const orc_edges *e, const char *prefix, int n_files);
orc_edges *orc_edges_read(const char *prefix);
int orc_cmd_count(const char *read_lib_file, int k, int min_count, const char *out_prefix, int threads);

/* ---- seq2sdbg / read2sdbg -------------------------------------------- */
typedef struct orc_seqs orc_seqs;
orc_seqs *orc_seqs_new(void);
void orc_seqs_free(orc_seqs *s);
/* sequence in STORED orientation (what megahit holds in its SeqPackage). */
void orc_seqs_add(orc_seqs *s, const uint8_t *bases, int64_t len, int mult);
void orc_seqs_add_edges(orc_seqs *s, const orc_edges *e);
/* contig FASTA as megahit writes it; reversed on load (contig_reverse=true). */
int orc_seqs_add_contigs(orc_seqs *s, const char *path, int min_len, int extend_loop, int k_from, int k_to);
int64_t orc_seqs_count(const orc_seqs *s);

typedef struct orc_sdbg orc_sdbg;
orc_sdbg *orc_seq2sdbg(const orc_seqs *s, int k, int threads);
orc_sdbg *orc_read2sdbg(const uint8_t *bases, const int64_t *starts, int64_t n_reads,
                        int k, int min_count, int threads);
void orc_sdbg_free(orc_sdbg *g);
int orc_sdbg_k(const orc_sdbg *g);
int orc_sdbg_words_per_tip(const orc_sdbg *g);
int64_t orc_sdbg_n(const orc_sdbg *g);
int64_t orc_sdbg_n_tips(const orc_sdbg *g);
int64_t orc_sdbg_n_large(const orc_sdbg *g);
const uint8_t *orc_sdbg_w(const orc_sdbg *g);
const uint8_t *orc_sdbg_last(const orc_sdbg *g);
const uint8_t *orc_sdbg_tip(const orc_sdbg *g);
const uint16_t *orc_sdbg_mul(const orc_sdbg *g);
const uint32_t *orc_sdbg_tip_labels(const orc_sdbg *g);
const int64_t *orc_sdbg_bucket_items(const orc_sdbg *g); /* 65536 */
int orc_sdbg_write(const orc_sdbg *g, const char *prefix, int n_files);
orc_sdbg *orc_sdbg_read(const char *prefix);
int orc_cmd_seq2sdbg(int k, int k_from, const char *input_prefix, const char *contig, const char *bubble,
                     const char *addi_contig, const char *local_contig, const char *out_prefix, int threads);
int orc_cmd_read2sdbg(const char *read_lib_file, int k, int min_count, const char *out_prefix, int threads);

const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
