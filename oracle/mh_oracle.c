/*
 * oracle/mh_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See mh_oracle.h.
 *
 * PARITY UNPINNED (megahit v1.2.9 source/binary unavailable; restated from
 * recollection of the upstream files named next to each function).
 *
 * Conventions restated from megahit v1.2.9 src/definitions.h:
 *   2 bits per base, A=0 C=1 G=2 T=3, 16 bases per uint32, first base in the
 *   most significant bits; bucket = first 8 bases (65536 buckets); `$` = 4;
 *   multiplicity uint16 capped at 65535; sdbg small multiplicity <= 254 inline,
 *   255 = "large, follows as uint16".
 */
#define _GNU_SOURCE
#include "mh_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <errno.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[512];
const char *orc_last_error(void) { return g_err; }
static int fail(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return -1;
}
static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory (%zu bytes)\n", n); abort(); }
  return p;
}
static void *xrealloc(void *q, size_t n) {
  void *p = realloc(q, n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory (%zu bytes)\n", n); abort(); }
  return p;
}
static inline int div_ceil(int a, int b) { return (a + b - 1) / b; }

/* pack n chars (values 0..3) MSB-first into W words, zero padded. */
static inline void pack_chars(const uint8_t *c, int n, uint32_t *out, int W) {
  for (int i = 0; i < W; ++i) out[i] = 0;
  for (int i = 0; i < n; ++i) out[i >> 4] |= (uint32_t)c[i] << (30 - 2 * (i & 15));
}
static inline int cmp_words(const uint32_t *a, const uint32_t *b, int W) {
  for (int i = 0; i < W; ++i) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}

/* ======================================================================
 * generic record sort: n records of W uint32 words, ascending by words.
 * Shape follows megahit's BaseSequenceSortingEngine (src/sorting/base_engine.cpp):
 * bucket by the first 8 bases (top 16 bits), then radix-sort each bucket.
 * ====================================================================== */
static void insertion_sort(uint32_t *a, int64_t n, int W) {
  uint32_t tmp[64];
  for (int64_t i = 1; i < n; ++i) {
    if (cmp_words(a + (i - 1) * W, a + i * W, W) <= 0) continue;
    memcpy(tmp, a + i * W, 4 * W);
    int64_t j = i;
    while (j > 0 && cmp_words(a + (j - 1) * W, tmp, W) > 0) {
      memcpy(a + j * W, a + (j - 1) * W, 4 * W);
      --j;
    }
    memcpy(a + j * W, tmp, 4 * W);
  }
}
/* sort one bucket living in src (n records); result must end in dst_final. other = scratch of same size */
static void radix_bucket(uint32_t *in_tmp, uint32_t *in_a, int64_t n, int W) {
  /* data starts in in_tmp, must end in in_a */
  if (n <= 48) {
    memcpy(in_a, in_tmp, (size_t)n * W * 4);
    insertion_sort(in_a, n, W);
    return;
  }
  uint32_t *src = in_tmp, *dst = in_a;
  for (int w = W - 1; w >= 0; --w) {
    int nbytes = (w == 0) ? 2 : 4;
    for (int b = 0; b < nbytes; ++b) {
      int64_t cnt[257];
      memset(cnt, 0, sizeof cnt);
      int sh = 8 * b;
      for (int64_t i = 0; i < n; ++i) cnt[((src[i * W + w] >> sh) & 255) + 1]++;
      int single = 0;
      for (int d = 0; d < 256; ++d) if (cnt[d + 1] == n) { single = 1; break; }
      if (single) continue;
      for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
      for (int64_t i = 0; i < n; ++i) {
        int64_t p = cnt[(src[i * W + w] >> sh) & 255]++;
        memcpy(dst + p * W, src + i * W, 4 * W);
      }
      uint32_t *t = src; src = dst; dst = t;
    }
  }
  if (src != in_a) memcpy(in_a, src, (size_t)n * W * 4);
}
static void sort_records(uint32_t *a, int64_t n, int W, int threads, int64_t *bucket_start /* 65537 or NULL */) {
  if (threads < 1) threads = 1;
  uint32_t *tmp = xmalloc((size_t)n * W * 4);
  int64_t *hist = xmalloc(sizeof(int64_t) * (size_t)threads * ORC_NUM_BUCKETS);
  memset(hist, 0, sizeof(int64_t) * (size_t)threads * ORC_NUM_BUCKETS);
  int64_t *bstart = xmalloc(sizeof(int64_t) * (ORC_NUM_BUCKETS + 1));
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    int t = 0, nt = 1;
#endif
    int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    int64_t *h = hist + (size_t)t * ORC_NUM_BUCKETS;
    for (int64_t i = lo; i < hi; ++i) h[a[i * W] >> 16]++;
#pragma omp barrier
#pragma omp single
    {
      int64_t acc = 0;
      for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
        bstart[b] = acc;
        for (int tt = 0; tt < nt; ++tt) {
          int64_t c = hist[(size_t)tt * ORC_NUM_BUCKETS + b];
          hist[(size_t)tt * ORC_NUM_BUCKETS + b] = acc;
          acc += c;
        }
      }
      bstart[ORC_NUM_BUCKETS] = acc;
    }
    for (int64_t i = lo; i < hi; ++i) {
      int64_t p = h[a[i * W] >> 16]++;
      memcpy(tmp + p * W, a + i * W, 4 * W);
    }
  }
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
  for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
    int64_t s = bstart[b], e = bstart[b + 1];
    if (e > s) radix_bucket(tmp + s * W, a + s * W, e - s, W);
  }
  if (bucket_start) memcpy(bucket_start, bstart, sizeof(int64_t) * (ORC_NUM_BUCKETS + 1));
  free(bstart); free(hist); free(tmp);
}

/* ======================================================================
 * reads  (megahit v1.2.9 src/sequence/io/{fastx_reader,binary_writer,binary_reader}.h,
 *         src/sequence/sequence_package.h, src/sequence/lib_io / lib_info)
 * ====================================================================== */
struct orc_reads {
  int64_t n, cap_n, nb, cap_b;
  uint8_t *bases;
  int64_t *starts;
  int max_len;
};
orc_reads *orc_reads_new(void) {
  orc_reads *r = xmalloc(sizeof *r);
  memset(r, 0, sizeof *r);
  r->cap_n = 1024; r->starts = xmalloc(sizeof(int64_t) * (r->cap_n + 1)); r->starts[0] = 0;
  r->cap_b = 1 << 16; r->bases = xmalloc(r->cap_b);
  return r;
}
void orc_reads_free(orc_reads *r) { if (r) { free(r->bases); free(r->starts); free(r); } }
int64_t orc_reads_count(const orc_reads *r) { return r->n; }
int64_t orc_reads_bases(const orc_reads *r) { return r->nb; }
int orc_reads_max_len(const orc_reads *r) { return r->max_len; }
const uint8_t *orc_reads_data(const orc_reads *r) { return r->bases; }
const int64_t *orc_reads_starts(const orc_reads *r) { return r->starts; }

/* SequencePackage dna_map: "ACGTNacgtn" -> "0123201232"; anything else -> 0. */
static inline uint8_t dna_code(char ch) {
  switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 'N': case 'n': return 2;
    default: return 0;
  }
}
static void reads_push(orc_reads *r, const char *s, int64_t len) {
  if (r->n + 1 > r->cap_n) { r->cap_n *= 2; r->starts = xrealloc(r->starts, sizeof(int64_t) * (r->cap_n + 1)); }
  while (r->nb + len > r->cap_b) { r->cap_b *= 2; r->bases = xrealloc(r->bases, r->cap_b); }
  for (int64_t i = 0; i < len; ++i) r->bases[r->nb + i] = dna_code(s[i]);
  r->nb += len;
  r->n += 1;
  r->starts[r->n] = r->nb;
  if (len > r->max_len) r->max_len = (int)len;
}
/* FastxReader::TrimN: [first non-N, first N after it). */
void orc_reads_add_ascii(orc_reads *r, const char *s, int64_t len, int n_policy) {
  if (n_policy == ORC_N_SPLIT) {
    int64_t i = 0; int pushed = 0;
    while (i < len) {
      while (i < len && (s[i] == 'N' || s[i] == 'n')) ++i;
      int64_t b = i;
      while (i < len && !(s[i] == 'N' || s[i] == 'n')) ++i;
      if (i > b) { reads_push(r, s + b, i - b); pushed = 1; }
    }
    if (!pushed) reads_push(r, s, 0);
    return;
  }
  int64_t b = len, e, i;
  for (i = 0; i < len; ++i) {
    if (s[i] == 'N' || s[i] == 'n') {
      if (b < len) break;
    } else if (b == len) {
      b = i;
    }
  }
  e = i;
  if (b > e) b = e;
  reads_push(r, s + b, e - b);
}

/* --- minimal kseq-like FASTA/FASTQ record iterator over an in-memory file --- */
typedef struct { char *buf; int64_t len, pos; } fx_file;
typedef struct { const char *name; int name_len; const char *comment; int comment_len; char *seq; int64_t seq_len; } fx_rec;
static int fx_open(fx_file *f, const char *path) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return fail("cannot open %s: %s", path, strerror(errno));
  int64_t cap = 1 << 20, len = 0; char *buf = xmalloc(cap);
  for (;;) {
    if (len == cap) { cap *= 2; buf = xrealloc(buf, cap); }
    size_t got = fread(buf + len, 1, cap - len, fp);
    if (got == 0) break;
    len += got;
  }
  fclose(fp);
  f->buf = buf; f->len = len; f->pos = 0;
  return 0;
}
static void fx_close(fx_file *f) { free(f->buf); f->buf = NULL; }
/* returns 1 if a record was read, 0 at EOF. seq is compacted in place (newlines removed). */
static int fx_next(fx_file *f, fx_rec *r) {
  char *b = f->buf; int64_t n = f->len, p = f->pos;
  while (p < n && b[p] != '>' && b[p] != '@') { while (p < n && b[p] != '\n') ++p; if (p < n) ++p; }
  if (p >= n) { f->pos = p; return 0; }
  int is_fq = b[p] == '@';
  ++p;
  int64_t h = p;
  while (p < n && b[p] != '\n') ++p;
  int64_t hend = p; if (hend > h && b[hend - 1] == '\r') --hend;
  int64_t q = h; while (q < hend && b[q] != ' ' && b[q] != '\t') ++q;
  r->name = b + h; r->name_len = (int)(q - h);
  while (q < hend && (b[q] == ' ' || b[q] == '\t')) ++q;
  r->comment = b + q; r->comment_len = (int)(hend - q);
  if (p < n) ++p;
  char *seq = b + p; int64_t sl = 0;
  while (p < n) {
    if (b[p] == '+' || b[p] == '>' || (b[p] == '@' && !is_fq)) break;
    if (b[p] == '@' && is_fq) break;
    while (p < n && b[p] != '\n') { if (b[p] != '\r') seq[sl++] = b[p]; ++p; }
    if (p < n) ++p;
  }
  r->seq = seq; r->seq_len = sl;
  if (p < n && b[p] == '+') {
    while (p < n && b[p] != '\n') ++p;
    if (p < n) ++p;
    int64_t ql = 0;
    while (p < n && ql < sl) {
      while (p < n && b[p] != '\n') { if (b[p] != '\r') ++ql; ++p; }
      if (p < n) ++p;
    }
  }
  f->pos = p;
  return 1;
}
int orc_reads_add_fastx(orc_reads *r, const char *path, int n_policy) {
  fx_file f; fx_rec rec;
  if (fx_open(&f, path)) return -1;
  while (fx_next(&f, &rec)) orc_reads_add_ascii(r, rec.seq, rec.seq_len, n_policy);
  fx_close(&f);
  return 0;
}
/* PairedFastxReader: mates alternate r1, r2, r1, r2 ... */
int orc_reads_add_fastx_pe(orc_reads *r, const char *p1, const char *p2, int n_policy) {
  fx_file f1, f2; fx_rec a, b;
  if (fx_open(&f1, p1)) return -1;
  if (fx_open(&f2, p2)) { fx_close(&f1); return -1; }
  int rc = 0;
  for (;;) {
    int g1 = fx_next(&f1, &a), g2 = fx_next(&f2, &b);
    if (!g1 && !g2) break;
    if (g1 != g2) { rc = fail("paired files have different numbers of reads"); break; }
    orc_reads_add_ascii(r, a.seq, a.seq_len, n_policy);
    orc_reads_add_ascii(r, b.seq, b.seq_len, n_policy);
  }
  fx_close(&f1); fx_close(&f2);
  return rc;
}
/* BinaryWriter: per read uint32 length then ceil(len/16) packed words. */
int orc_reads_write_bin(const orc_reads *r, const char *bin_path) {
  FILE *fp = fopen(bin_path, "wb");
  if (!fp) return fail("cannot create %s: %s", bin_path, strerror(errno));
  uint32_t *w = xmalloc(4 * (size_t)(r->max_len / 16 + 2));
  for (int64_t i = 0; i < r->n; ++i) {
    uint32_t len = (uint32_t)(r->starts[i + 1] - r->starts[i]);
    int nw = div_ceil((int)len, 16);
    pack_chars(r->bases + r->starts[i], (int)len, w, nw);
    fwrite(&len, 4, 1, fp);
    fwrite(w, 4, nw, fp);
  }
  free(w);
  fclose(fp);
  return 0;
}
orc_reads *orc_reads_load_bin(const char *bin_path) {
  FILE *fp = fopen(bin_path, "rb");
  if (!fp) { fail("cannot open %s: %s", bin_path, strerror(errno)); return NULL; }
  orc_reads *r = orc_reads_new();
  uint32_t len; uint32_t *w = NULL; char *s = NULL; size_t cap = 0;
  while (fread(&len, 4, 1, fp) == 1) {
    size_t nw = (len + 15) / 16;
    if (nw + 1 > cap) { cap = nw + 64; w = xrealloc(w, 4 * cap); s = xrealloc(s, 16 * cap); }
    if (fread(w, 4, nw, fp) != nw) { fail("truncated %s", bin_path); orc_reads_free(r); r = NULL; break; }
    for (uint32_t i = 0; i < len; ++i) s[i] = "ACGT"[(w[i >> 4] >> (30 - 2 * (i & 15))) & 3];
    reads_push(r, s, len);
  }
  free(w); free(s); fclose(fp);
  return r;
}
static char *read_line(FILE *fp) {
  size_t cap = 256, n = 0; char *s = xmalloc(cap); int c;
  while ((c = fgetc(fp)) != EOF && c != '\n') { if (n + 2 > cap) { cap *= 2; s = xrealloc(s, cap); } s[n++] = (char)c; }
  if (c == EOF && n == 0) { free(s); return NULL; }
  if (n && s[n - 1] == '\r') --n;
  s[n] = 0;
  return s;
}
/* megahit_core buildlib: lib file = repeated (metadata line, "pe f1 f2" | "se f" | "interleaved f").
 * reads.lib layout written by /root/reference/assemble/assemble_wrapper.py:166-190.
 * .lib_info: "<bases> <reads>\n" then per library "<metadata>\n" "<pe|se> <from> <to> <max_len>\n". */
int orc_cmd_buildlib(const char *lib_file, const char *out_prefix, int n_policy) {
  FILE *fp = fopen(lib_file, "r");
  if (!fp) return fail("cannot open %s: %s", lib_file, strerror(errno));
  orc_reads *all = orc_reads_new();
  char info_path[4096], bin_path[4096];
  snprintf(info_path, sizeof info_path, "%s.lib_info", out_prefix);
  snprintf(bin_path, sizeof bin_path, "%s.bin", out_prefix);
  size_t lcap = 4096, llen = 0; char *libtxt = xmalloc(lcap); libtxt[0] = 0;
  int rc = 0; char *meta;
  while ((meta = read_line(fp)) != NULL) {
    char *spec = read_line(fp);
    if (!spec) { free(meta); break; }
    char type[32] = {0}, f1[2048] = {0}, f2[2048] = {0};
    int nt = sscanf(spec, "%31s %2047s %2047s", type, f1, f2);
    int64_t from = all->n; int paired = 0; int lib_max = 0;
    int64_t before_max = all->max_len; all->max_len = 0;
    if (nt >= 3 && !strcmp(type, "pe")) { paired = 1; rc = orc_reads_add_fastx_pe(all, f1, f2, n_policy); }
    else if (nt >= 2 && !strcmp(type, "se")) rc = orc_reads_add_fastx(all, f1, n_policy);
    else if (nt >= 2 && !strcmp(type, "interleaved")) { paired = 1; rc = orc_reads_add_fastx(all, f1, n_policy); }
    else rc = fail("bad library line: %s", spec);
    lib_max = all->max_len; if (before_max > all->max_len) all->max_len = (int)before_max;
    if (!rc) {
      char line[8192];
      int m = snprintf(line, sizeof line, "%s\n%s %lld %lld %d\n", meta, paired ? "pe" : "se",
                       (long long)from, (long long)all->n - 1, lib_max);
      while (llen + m + 1 > lcap) { lcap *= 2; libtxt = xrealloc(libtxt, lcap); }
      memcpy(libtxt + llen, line, m + 1); llen += m;
    }
    free(meta); free(spec);
    if (rc) break;
  }
  fclose(fp);
  if (!rc) rc = orc_reads_write_bin(all, bin_path);
  if (!rc) {
    FILE *fi = fopen(info_path, "w");
    if (!fi) rc = fail("cannot create %s", info_path);
    else { fprintf(fi, "%lld %lld\n%s", (long long)all->nb, (long long)all->n, libtxt); fclose(fi); }
  }
  free(libtxt);
  orc_reads_free(all);
  return rc;
}

/* ======================================================================
 * count  (megahit v1.2.9 src/sorting/kmer_counter.cpp)
 * ====================================================================== */
struct orc_edges {
  int k, words, sorted;
  int64_t n;
  uint32_t *data;
  int64_t bucket_counts[ORC_NUM_BUCKETS];
  int64_t counting[ORC_MAX_MUL + 1];
};
void orc_edges_free(orc_edges *e) { if (e) { free(e->data); free(e); } }
int orc_edges_k(const orc_edges *e) { return e->k; }
int orc_edges_words(const orc_edges *e) { return e->words; }
int orc_edges_sorted(const orc_edges *e) { return e->sorted; }
int64_t orc_edges_n(const orc_edges *e) { return e->n; }
const uint32_t *orc_edges_data(const orc_edges *e) { return e->data; }
const int64_t *orc_edges_bucket_counts(const orc_edges *e) { return e->bucket_counts; }
const int64_t *orc_edges_counting(const orc_edges *e) { return e->counting; }

/* Canonical (k+1)-mer of the REVERSED read at every position
 * (KmerCounter::Initialize reads the library with is_reverse = true;
 *  Lv0CalcBucketSize / Lv1FillOffsets: strand 1 iff rev_edge < edge, ties -> strand 0;
 *  Lv2ExtractSubString copies the chosen strand's k+1 bases). */
static int64_t gen_count_keys(const uint8_t *bases, const int64_t *starts, int64_t n_reads, int k, int W,
                              int threads, uint32_t **out) {
  int64_t *off = xmalloc(sizeof(int64_t) * (n_reads + 1));
  off[0] = 0;
  for (int64_t r = 0; r < n_reads; ++r) {
    int64_t L = starts[r + 1] - starts[r];
    off[r + 1] = off[r] + (L >= k + 1 ? L - k : 0);
  }
  int64_t n = off[n_reads];
  uint32_t *keys = xmalloc((size_t)n * W * 4);
#pragma omp parallel num_threads(threads)
  {
    int cap = 1024; uint8_t *s = xmalloc(cap), *rc = xmalloc(cap);
    uint32_t a[64], b[64];
#pragma omp for schedule(dynamic, 4096)
    for (int64_t r = 0; r < n_reads; ++r) {
      int64_t L = starts[r + 1] - starts[r];
      if (L < k + 1) continue;
      if (L > cap) { cap = (int)L * 2; s = xrealloc(s, cap); rc = xrealloc(rc, cap); }
      const uint8_t *c = bases + starts[r];
      for (int64_t j = 0; j < L; ++j) { s[j] = c[L - 1 - j]; }          /* stored = reversed read   */
      for (int64_t j = 0; j < L; ++j) { rc[j] = 3 - s[L - 1 - j]; }      /* revcomp of stored string */
      uint32_t *o = keys + off[r] * W;
      for (int64_t i = 0; i + k + 1 <= L; ++i) {
        pack_chars(s + i, k + 1, a, W);
        pack_chars(rc + (L - 1 - i - k), k + 1, b, W);
        memcpy(o, cmp_words(b, a, W) < 0 ? b : a, 4 * W);
        o += W;
      }
    }
    free(s); free(rc);
  }
  free(off);
  *out = keys;
  return n;
}
/* KmerCounter::PackEdge */
static inline void pack_edge(uint32_t *dst, const uint32_t *key, int k, int Wk, int We, int64_t count) {
  for (int i = 0; i < We; ++i) dst[i] = i < Wk ? key[i] : 0;
  int chars_in_last = (k + 1) % 16, which = (k + 1) / 16;
  if (chars_in_last > 0) {
    dst[which] >>= (16 - chars_in_last) * 2;
    dst[which] <<= (16 - chars_in_last) * 2;
  } else if (which < We) {
    dst[which] = 0;
  }
  for (int i = which + 1; i < We; ++i) dst[i] = 0;
  dst[We - 1] |= (uint32_t)(count > ORC_MAX_MUL ? ORC_MAX_MUL : count);
}
orc_edges *orc_count(const uint8_t *bases, const int64_t *starts, int64_t n_reads, int k, int min_count, int threads) {
  if (k < 9 || k > 255) { fail("k out of range"); return NULL; }
  if (threads < 1) threads = 1;
  int Wk = div_ceil(2 * (k + 1), 32), We = div_ceil(2 * (k + 1) + 16, 32);
  uint32_t *keys; int64_t n = gen_count_keys(bases, starts, n_reads, k, Wk, threads, &keys);
  int64_t *bstart = xmalloc(sizeof(int64_t) * (ORC_NUM_BUCKETS + 1));
  sort_records(keys, n, Wk, threads, bstart);
  orc_edges *e = xmalloc(sizeof *e);
  memset(e, 0, sizeof *e);
  e->k = k; e->words = We; e->sorted = 1;
  /* KmerCounter::Lv2Postprocess: runs of equal keys -> count, histogram, solid filter.  Runs never straddle a
   * bucket, so buckets are processed independently (first pass counts, second pass writes). */
  int64_t *hist = xmalloc(sizeof(int64_t) * (size_t)threads * (ORC_MAX_MUL + 1));
  memset(hist, 0, sizeof(int64_t) * (size_t)threads * (ORC_MAX_MUL + 1));
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
  for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
#ifdef _OPENMP
    int64_t *h = hist + (size_t)omp_get_thread_num() * (ORC_MAX_MUL + 1);
#else
    int64_t *h = hist;
#endif
    int64_t solid = 0;
    for (int64_t i = bstart[b], j; i < bstart[b + 1]; i = j) {
      j = i + 1;
      while (j < bstart[b + 1] && cmp_words(keys + i * Wk, keys + j * Wk, Wk) == 0) ++j;
      int64_t c = j - i;
      h[c > ORC_MAX_MUL ? ORC_MAX_MUL : c]++;
      solid += c >= min_count;
    }
    e->bucket_counts[b] = solid;
  }
  for (int t = 0; t < threads; ++t)
    for (int c = 0; c <= ORC_MAX_MUL; ++c) e->counting[c] += hist[(size_t)t * (ORC_MAX_MUL + 1) + c];
  free(hist);
  int64_t *eoff = xmalloc(sizeof(int64_t) * (ORC_NUM_BUCKETS + 1));
  eoff[0] = 0;
  for (int b = 0; b < ORC_NUM_BUCKETS; ++b) eoff[b + 1] = eoff[b] + e->bucket_counts[b];
  e->n = eoff[ORC_NUM_BUCKETS];
  e->data = xmalloc((size_t)e->n * We * 4);
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
  for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
    int64_t o = eoff[b];
    for (int64_t i = bstart[b], j; i < bstart[b + 1]; i = j) {
      j = i + 1;
      while (j < bstart[b + 1] && cmp_words(keys + i * Wk, keys + j * Wk, Wk) == 0) ++j;
      if (j - i >= min_count) pack_edge(e->data + (o++) * We, keys + i * Wk, k, Wk, We, j - i);
    }
  }
  free(eoff); free(bstart);
  free(keys);
  return e;
}
/* EdgeWriter / EdgeIoMetadata::Serialize (src/sequence/io/edge/edge_writer.h, edge_io_meta.h) */
int orc_edges_write(const orc_edges *e, const char *prefix, int n_files) {
  if (n_files < 1) n_files = 1;
  char path[4096];
  FILE **fps = xmalloc(sizeof(FILE *) * n_files);
  for (int f = 0; f < n_files; ++f) {
    snprintf(path, sizeof path, "%s.edges.%d", prefix, f);
    fps[f] = fopen(path, "wb");
    if (!fps[f]) return fail("cannot create %s", path);
  }
  snprintf(path, sizeof path, "%s.edges.info", prefix);
  FILE *fi = fopen(path, "w");
  if (!fi) return fail("cannot create %s", path);
  fprintf(fi, "kmer_size %d\nwords_per_edge %d\nnum_files %d\nnum_buckets %d\nnum_edges %lld\nis_sorted %d\n",
          e->k, e->words, n_files, e->sorted ? ORC_NUM_BUCKETS : 0, (long long)e->n, e->sorted);
  if (e->sorted) {
    int64_t pos = 0; int64_t *foff = xmalloc(sizeof(int64_t) * n_files);
    memset(foff, 0, sizeof(int64_t) * n_files);
    for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
      int f = (int)((int64_t)b * n_files / ORC_NUM_BUCKETS);
      int64_t c = e->bucket_counts[b];
      if (c) {
        fwrite(e->data + pos * e->words, 4, (size_t)c * e->words, fps[f]);
        fprintf(fi, "%d %d %lld %lld\n", b, f, (long long)foff[f], (long long)c);
        foff[f] += c; pos += c;
      } else {
        fprintf(fi, "%d -1 0 0\n", b);
      }
    }
    free(foff);
  } else {
    fwrite(e->data, 4, (size_t)e->n * e->words, fps[0]);
  }
  fclose(fi);
  for (int f = 0; f < n_files; ++f) fclose(fps[f]);
  free(fps);
  return 0;
}
static int scan_field(FILE *fp, const char *name, long long *v) {
  char tok[64];
  if (fscanf(fp, "%63s %lld", tok, v) != 2 || strcmp(tok, name)) return fail("bad field, expected %s", name);
  return 0;
}
orc_edges *orc_edges_read(const char *prefix) {
  char path[4096];
  snprintf(path, sizeof path, "%s.edges.info", prefix);
  FILE *fi = fopen(path, "r");
  if (!fi) { fail("cannot open %s", path); return NULL; }
  long long k, words, nfiles, nbuckets, nedges, sorted;
  if (scan_field(fi, "kmer_size", &k) || scan_field(fi, "words_per_edge", &words) || scan_field(fi, "num_files", &nfiles) ||
      scan_field(fi, "num_buckets", &nbuckets) || scan_field(fi, "num_edges", &nedges) || scan_field(fi, "is_sorted", &sorted)) {
    fclose(fi); return NULL;
  }
  orc_edges *e = xmalloc(sizeof *e);
  memset(e, 0, sizeof *e);
  e->k = (int)k; e->words = (int)words; e->sorted = (int)sorted; e->n = nedges;
  e->data = xmalloc((size_t)nedges * words * 4);
  uint32_t **fdata = xmalloc(sizeof(uint32_t *) * nfiles);
  int64_t *fcount = xmalloc(sizeof(int64_t) * nfiles);
  for (int f = 0; f < nfiles; ++f) {
    snprintf(path, sizeof path, "%s.edges.%d", prefix, f);
    FILE *fp = fopen(path, "rb");
    if (!fp) { fail("cannot open %s", path); return NULL; }
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    fdata[f] = xmalloc(sz); fcount[f] = sz / (4 * words);
    if (fread(fdata[f], 1, sz, fp) != (size_t)sz) { fail("short read %s", path); return NULL; }
    fclose(fp);
  }
  int64_t pos = 0;
  if (sorted) {
    for (int b = 0; b < nbuckets; ++b) {
      long long bid, fid, off, cnt;
      if (fscanf(fi, "%lld %lld %lld %lld", &bid, &fid, &off, &cnt) != 4 || bid != b) { fail("bucket id not matched"); return NULL; }
      if (fid < 0 || cnt == 0) continue;
      memcpy(e->data + pos * words, fdata[fid] + off * words, (size_t)cnt * words * 4);
      e->bucket_counts[b] = cnt; pos += cnt;
    }
  } else {
    for (int f = 0; f < nfiles; ++f) { memcpy(e->data + pos * words, fdata[f], (size_t)fcount[f] * words * 4); pos += fcount[f]; }
  }
  fclose(fi);
  for (int f = 0; f < nfiles; ++f) free(fdata[f]);
  free(fdata); free(fcount);
  if (pos != nedges) { fail("edge count mismatch %lld vs %lld", (long long)pos, nedges); orc_edges_free(e); return NULL; }
  return e;
}
static orc_reads *load_lib(const char *read_lib_file) {
  char path[4096];
  snprintf(path, sizeof path, "%s.bin", read_lib_file);
  return orc_reads_load_bin(path);
}
int orc_cmd_count(const char *read_lib_file, int k, int min_count, const char *out_prefix, int threads) {
  orc_reads *r = load_lib(read_lib_file);
  if (!r) return -1;
  orc_edges *e = orc_count(r->bases, r->starts, r->n, k, min_count, threads);
  orc_reads_free(r);
  if (!e) return -1;
  int rc = orc_edges_write(e, out_prefix, threads);
  if (!rc) {
    /* Lv0Postprocess: "<prefix>.counting", cumulative distinct-edge histogram */
    char path[4096]; snprintf(path, sizeof path, "%s.counting", out_prefix);
    FILE *fc = fopen(path, "w");
    if (fc) {
      long long acc = 0;
      for (int i = 1; i <= ORC_MAX_MUL; ++i) { acc += e->counting[i]; fprintf(fc, "%d %lld\n", i, acc); }
      fclose(fc);
    }
  }
  orc_edges_free(e);
  return rc;
}

/* ======================================================================
 * seq2sdbg  (megahit v1.2.9 src/sorting/seq_to_sdbg.cpp)
 * ====================================================================== */
struct orc_seqs {
  int64_t n, cap_n, nb, cap_b;
  uint8_t *bases; int64_t *starts; int *mult;
};
orc_seqs *orc_seqs_new(void) {
  orc_seqs *s = xmalloc(sizeof *s);
  memset(s, 0, sizeof *s);
  s->cap_n = 1024; s->starts = xmalloc(sizeof(int64_t) * (s->cap_n + 1)); s->starts[0] = 0;
  s->mult = xmalloc(sizeof(int) * s->cap_n);
  s->cap_b = 1 << 16; s->bases = xmalloc(s->cap_b);
  return s;
}
void orc_seqs_free(orc_seqs *s) { if (s) { free(s->bases); free(s->starts); free(s->mult); free(s); } }
int64_t orc_seqs_count(const orc_seqs *s) { return s->n; }
void orc_seqs_add(orc_seqs *s, const uint8_t *b, int64_t len, int mult) {
  if (s->n + 1 > s->cap_n) {
    s->cap_n *= 2;
    s->starts = xrealloc(s->starts, sizeof(int64_t) * (s->cap_n + 1));
    s->mult = xrealloc(s->mult, sizeof(int) * s->cap_n);
  }
  while (s->nb + len > s->cap_b) { s->cap_b *= 2; s->bases = xrealloc(s->bases, s->cap_b); }
  memcpy(s->bases + s->nb, b, len);
  s->nb += len;
  s->mult[s->n] = mult;
  s->n++;
  s->starts[s->n] = s->nb;
}
/* SeqToSdbg::Initialize: every edge record is a (k+1)-base sequence with its multiplicity. */
void orc_seqs_add_edges(orc_seqs *s, const orc_edges *e) {
  int K1 = e->k + 1; uint8_t buf[512];
  for (int64_t i = 0; i < e->n; ++i) {
    const uint32_t *w = e->data + i * e->words;
    for (int j = 0; j < K1; ++j) buf[j] = (w[j >> 4] >> (30 - 2 * (j & 15))) & 3;
    orc_seqs_add(s, buf, K1, (int)(w[e->words - 1] & ORC_MAX_MUL));
  }
}
/* ContigReader (+SetExtendLoop(k_from,k_to), SetMinLen) with contig_reverse = true.
 * Header grammar ">k<K>_<id> flag=<f> multi=<float> len=<n>" (same grammar
 * /root/reference/assemble/fastfilter_src/src/main.rs:81-101 parses); flag bit 1 = standalone, bit 2 = loop;
 * multiplicity = min(65535, int(multi + 0.5)). */
int orc_seqs_add_contigs(orc_seqs *s, const char *path, int min_len, int extend_loop, int k_from, int k_to) {
  fx_file f; fx_rec rec;
  if (fx_open(&f, path)) return -1;
  uint8_t *buf = NULL; int64_t cap = 0;
  while (fx_next(&f, &rec)) {
    unsigned flag = 0; float multi = 1.0f; char cm[256];
    int cl = rec.comment_len < 255 ? rec.comment_len : 255;
    memcpy(cm, rec.comment, cl); cm[cl] = 0;
    sscanf(cm, "flag=%u multi=%f", &flag, &multi);
    int64_t L = rec.seq_len, ext = 0;
    if (extend_loop && (flag & 2u)) ext = k_to - k_from;
    if (ext > L) ext = L;
    if (L + ext < min_len) continue;
    if (L + ext > cap) { cap = (L + ext) * 2; buf = xrealloc(buf, cap); }
    for (int64_t i = 0; i < L; ++i) buf[i] = dna_code(rec.seq[i]);
    for (int64_t i = 0; i < ext; ++i) buf[L + i] = buf[i];
    L += ext;
    for (int64_t i = 0; i < L / 2; ++i) { uint8_t t = buf[i]; buf[i] = buf[L - 1 - i]; buf[L - 1 - i] = t; }
    int m = (int)(multi + 0.5f);
    if (m > ORC_MAX_MUL) m = ORC_MAX_MUL;
    orc_seqs_add(s, buf, L, m);
  }
  free(buf);
  fx_close(&f);
  return 0;
}

struct orc_sdbg {
  int k, words_per_tip;
  int64_t n, n_tips, n_large, cap, cap_tips;
  uint8_t *w, *last, *tip; uint16_t *mul; uint32_t *tip_labels;
  int64_t bucket_items[ORC_NUM_BUCKETS], bucket_tips[ORC_NUM_BUCKETS], bucket_large[ORC_NUM_BUCKETS];
};
void orc_sdbg_free(orc_sdbg *g) { if (g) { free(g->w); free(g->last); free(g->tip); free(g->mul); free(g->tip_labels); free(g); } }
int orc_sdbg_k(const orc_sdbg *g) { return g->k; }
int orc_sdbg_words_per_tip(const orc_sdbg *g) { return g->words_per_tip; }
int64_t orc_sdbg_n(const orc_sdbg *g) { return g->n; }
int64_t orc_sdbg_n_tips(const orc_sdbg *g) { return g->n_tips; }
int64_t orc_sdbg_n_large(const orc_sdbg *g) { return g->n_large; }
const uint8_t *orc_sdbg_w(const orc_sdbg *g) { return g->w; }
const uint8_t *orc_sdbg_last(const orc_sdbg *g) { return g->last; }
const uint8_t *orc_sdbg_tip(const orc_sdbg *g) { return g->tip; }
const uint16_t *orc_sdbg_mul(const orc_sdbg *g) { return g->mul; }
const uint32_t *orc_sdbg_tip_labels(const orc_sdbg *g) { return g->tip_labels; }
const int64_t *orc_sdbg_bucket_items(const orc_sdbg *g) { return g->bucket_items; }

static orc_sdbg *sdbg_new(int k, int64_t cap) {
  orc_sdbg *g = xmalloc(sizeof *g);
  memset(g, 0, sizeof *g);
  g->k = k; g->words_per_tip = div_ceil(2 * k, 32);
  g->cap = cap > 16 ? cap : 16; g->cap_tips = 1024;
  g->w = xmalloc(g->cap); g->last = xmalloc(g->cap); g->tip = xmalloc(g->cap);
  g->mul = xmalloc(2 * g->cap); g->tip_labels = xmalloc(4 * g->cap_tips * g->words_per_tip);
  return g;
}
/* SdbgWriter::Write bookkeeping */
static void sdbg_push(orc_sdbg *g, int bucket, int w, int last, int tip, int mul, const uint32_t *label) {
  if (g->n == g->cap) {
    g->cap *= 2;
    g->w = xrealloc(g->w, g->cap); g->last = xrealloc(g->last, g->cap); g->tip = xrealloc(g->tip, g->cap);
    g->mul = xrealloc(g->mul, 2 * g->cap);
  }
  g->w[g->n] = (uint8_t)w; g->last[g->n] = (uint8_t)last; g->tip[g->n] = (uint8_t)tip; g->mul[g->n] = (uint16_t)mul;
  g->n++;
  g->bucket_items[bucket]++;
  if (mul > 254) { g->n_large++; g->bucket_large[bucket]++; }
  if (tip) {
    if (g->n_tips == g->cap_tips) { g->cap_tips *= 2; g->tip_labels = xrealloc(g->tip_labels, 4 * g->cap_tips * g->words_per_tip); }
    memcpy(g->tip_labels + g->n_tips * g->words_per_tip, label, 4 * g->words_per_tip);
    g->n_tips++; g->bucket_tips[bucket]++;
  }
}

/* item layout helpers.  mode 0 = seq2sdbg (flag at bit 19, b at 16..18, 65535-mult in 0..15 of the last word);
 *                       mode 1 = read2sdbg stage 2 (flag at bit 3, b at 0..2). */
typedef struct { int k, W, mode; } item_fmt;
static inline int it_flag_shift(const item_fmt *f) { return f->mode == 0 ? 19 : 3; }
static inline int it_b_shift(const item_fmt *f) { return f->mode == 0 ? 16 : 0; }
static inline int it_a(const item_fmt *f, const uint32_t *it) {
  if (!((it[f->W - 1] >> it_flag_shift(f)) & 1)) return ORC_SENTINEL;
  int which = (f->k - 1) / 16, idx = (f->k - 1) % 16;
  return (it[which] >> (15 - idx) * 2) & 3;
}
static inline int it_b(const item_fmt *f, const uint32_t *it) { return (it[f->W - 1] >> it_b_shift(f)) & 7; }
/* IsDiffKMinusOneMer */
static inline int it_diff_km1(const item_fmt *f, const uint32_t *x, const uint32_t *y) {
  int chars_in_last = (f->k - 1) % 16, full = (f->k - 1) / 16;
  if (chars_in_last > 0) {
    int sh = (16 - chars_in_last) * 2;
    if ((x[full] >> sh) != (y[full] >> sh)) return 1;
  }
  for (int i = full - 1; i >= 0; --i) if (x[i] != y[i]) return 1;
  return 0;
}
/* SeqToSdbg::Lv2Postprocess / Read2SdbgS2::Lv2Postprocess over the sorted items [from, to) (whole (k-1)-prefix groups).
 * g == NULL: only tally into cnt[3] (items, tips, large); else write at item offset *io / tip offset *to_. */
static void sdbg_walk_range(const uint32_t *items, int64_t from, int64_t to, const item_fmt *f, orc_sdbg *g, int64_t *cnt,
                            int64_t *io, int64_t *to_) {
  int W = f->W, Wd = div_ceil(2 * f->k, 32);
  for (int64_t s = from, e; s < to; s = e) {
    e = s + 1;
    while (e < to && !it_diff_km1(f, items + s * W, items + e * W)) ++e;
    int has_solid_a = 0, has_solid_b = 0, outputed_b = 0;
    int64_t last_a[4] = {-1, -1, -1, -1};
    for (int64_t i = s; i < e; ++i) {
      int a = it_a(f, items + i * W), b = it_b(f, items + i * W);
      if (a != ORC_SENTINEL && b != ORC_SENTINEL) { has_solid_a |= 1 << a; has_solid_b |= 1 << b; }
      if (a != ORC_SENTINEL && (b != ORC_SENTINEL || !(has_solid_a & (1 << a)))) last_a[a] = i;
    }
    for (int64_t i = s, j; i < e; i = j) {
      const uint32_t *cur = items + i * W;
      int a = it_a(f, cur), b = it_b(f, cur);
      j = i + 1;
      while (j < e && it_a(f, items + j * W) == a && it_b(f, items + j * W) == b) ++j;
      int is_dollar = 0, mul;
      if (f->mode == 0) mul = ORC_MAX_MUL - (int)(cur[W - 1] & ORC_MAX_MUL);
      else mul = (int)((j - i) > ORC_MAX_MUL ? ORC_MAX_MUL : (j - i));
      if (a == ORC_SENTINEL) {
        if (has_solid_b & (1 << b)) continue;
        is_dollar = 1;
        if (f->mode == 1) mul = 0;
      }
      if (b == ORC_SENTINEL) {
        if (has_solid_a & (1 << a)) continue;
        if (f->mode == 1) mul = 0;
      }
      int w = (b == ORC_SENTINEL) ? 0 : ((outputed_b & (1 << b)) ? b + 5 : b + 1);
      outputed_b |= 1 << b;
      int last = (a == ORC_SENTINEL) ? 0 : (last_a[a] == j - 1 ? 1 : 0);
      if (!g) {
        cnt[0]++; cnt[1] += is_dollar; cnt[2] += mul > 254;
      } else {
        int64_t o = (*io)++;
        g->w[o] = (uint8_t)w; g->last[o] = (uint8_t)last; g->tip[o] = (uint8_t)is_dollar; g->mul[o] = (uint16_t)mul;
        if (is_dollar) memcpy(g->tip_labels + ((*to_)++) * Wd, cur, 4 * Wd);   /* raw first words of the item */
      }
    }
  }
}
static orc_sdbg *sdbg_postprocess(const uint32_t *items, int64_t n, const item_fmt *f, const int64_t *bstart, int threads) {
  orc_sdbg *g = sdbg_new(f->k, 16);
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
  for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
    int64_t cnt[3] = {0, 0, 0};
    if (bstart[b + 1] > bstart[b]) sdbg_walk_range(items, bstart[b], bstart[b + 1], f, NULL, cnt, NULL, NULL);
    g->bucket_items[b] = cnt[0]; g->bucket_tips[b] = cnt[1]; g->bucket_large[b] = cnt[2];
  }
  int64_t *io = xmalloc(sizeof(int64_t) * (ORC_NUM_BUCKETS + 1)), *to = xmalloc(sizeof(int64_t) * (ORC_NUM_BUCKETS + 1));
  io[0] = to[0] = 0;
  for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
    io[b + 1] = io[b] + g->bucket_items[b]; to[b + 1] = to[b] + g->bucket_tips[b];
    g->n_large += g->bucket_large[b];
  }
  g->n = io[ORC_NUM_BUCKETS]; g->n_tips = to[ORC_NUM_BUCKETS];
  free(g->w); free(g->last); free(g->tip); free(g->mul); free(g->tip_labels);
  g->cap = g->n + 1; g->cap_tips = g->n_tips + 1;
  g->w = xmalloc(g->cap); g->last = xmalloc(g->cap); g->tip = xmalloc(g->cap); g->mul = xmalloc(2 * g->cap);
  g->tip_labels = xmalloc(4 * g->cap_tips * g->words_per_tip);
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
  for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
    int64_t i0 = io[b], t0 = to[b];
    if (bstart[b + 1] > bstart[b]) sdbg_walk_range(items, bstart[b], bstart[b + 1], f, g, NULL, &i0, &t0);
  }
  free(io); free(to);
  (void)n;
  return g;
}
orc_sdbg *orc_seq2sdbg(const orc_seqs *s, int k, int threads) {
  if (k < 9 || k > 255) { fail("k out of range"); return NULL; }
  if (threads < 1) threads = 1;
  item_fmt f = {k, div_ceil(2 * k + 4 + 16, 32), 0};
  int W = f.W;
  int64_t *off = xmalloc(sizeof(int64_t) * (s->n + 1));
  off[0] = 0;
  for (int64_t i = 0; i < s->n; ++i) {
    int64_t L = s->starts[i + 1] - s->starts[i];
    off[i + 1] = off[i] + (L >= k + 1 ? 2 * (L - k + 2) : 0);
  }
  int64_t n = off[s->n];
  uint32_t *items = xmalloc((size_t)n * W * 4);
#pragma omp parallel num_threads(threads)
  {
    int64_t cap = 1024; uint8_t *rc = xmalloc(cap);
#pragma omp for schedule(dynamic, 1024)
    for (int64_t q = 0; q < s->n; ++q) {
      int64_t L = s->starts[q + 1] - s->starts[q];
      if (L < k + 1) continue;
      if (L > cap) { cap = L * 2; rc = xrealloc(rc, cap); }
      const uint8_t *fw = s->bases + s->starts[q];
      for (int64_t j = 0; j < L; ++j) rc[j] = 3 - fw[L - 1 - j];
      uint32_t *o = items + off[q] * W;
      for (int strand = 0; strand < 2; ++strand) {
        const uint8_t *t = strand ? rc : fw;
        /* Lv0CalcBucketSize loop bound: offsets 0 .. L-k+1 ("$xxxx, xxxxx, ..., xxxx$") */
        for (int64_t of = 0; of <= L - k + 1; ++of) {
          int nchars = (of + k > L) ? k - 1 : k;
          int prev = of == 0 ? ORC_SENTINEL : t[of - 1];
          int counting = (of > 0 && of + k <= L) ? s->mult[q] : 0;
          pack_chars(t + of, nchars, o, W);
          int inv = ORC_MAX_MUL - counting; if (inv < 0) inv = 0;
          o[W - 1] |= (uint32_t)(nchars == k) << 19;
          o[W - 1] |= (uint32_t)prev << 16;
          o[W - 1] |= (uint32_t)inv;      /* larger multiplicity sorts first */
          o += W;
        }
      }
    }
    free(rc);
  }
  free(off);
  int64_t *bstart = xmalloc(sizeof(int64_t) * (ORC_NUM_BUCKETS + 1));
  sort_records(items, n, W, threads, bstart);
  orc_sdbg *g = sdbg_postprocess(items, n, &f, bstart, threads);
  free(bstart);
  free(items);
  return g;
}

/* ======================================================================
 * read2sdbg  (megahit v1.2.9 src/sorting/read_to_sdbg_s1.cpp, read_to_sdbg_s2.cpp)
 * Stage 1's RESULT (the per-position solid bitmap) is restated through orc_count:
 * a position is solid iff its canonical (k+1)-mer occurs >= min_count times
 * (stage 1 is skipped when min_count <= 1).  Stage 2 is restated directly.
 * ====================================================================== */
static int edge_is_solid(const orc_edges *e, const uint32_t *key, int Wk) {
  int64_t lo = 0, hi = e->n;
  uint32_t probe[64];
  pack_edge(probe, key, e->k, Wk, e->words, 0);
  uint32_t tmp[64];
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    memcpy(tmp, e->data + mid * e->words, 4 * e->words);
    tmp[e->words - 1] &= ~(uint32_t)ORC_MAX_MUL;
    int c = cmp_words(tmp, probe, e->words);
    if (c == 0) return 1;
    if (c < 0) lo = mid + 1; else hi = mid;
  }
  return 0;
}
orc_sdbg *orc_read2sdbg(const uint8_t *bases, const int64_t *starts, int64_t n_reads, int k, int min_count, int threads) {
  if (k < 9 || k > 255) { fail("k out of range"); return NULL; }
  if (threads < 1) threads = 1;
  int Wk = div_ceil(2 * (k + 1), 32);
  orc_edges *solid = NULL;
  if (min_count > 1) { solid = orc_count(bases, starts, n_reads, k, min_count, threads); if (!solid) return NULL; }
  item_fmt f = {k, div_ceil(2 * k + 4, 32), 1};
  int W = f.W;
  /* pass 0: per-position solid flags (stage 1's bitmap); pass 1: item counts per read; pass 2: fill */
  int64_t tot_pos = 0;
  int64_t *pos_off = xmalloc(sizeof(int64_t) * (n_reads + 1));
  for (int64_t r = 0; r < n_reads; ++r) {
    int64_t L = starts[r + 1] - starts[r];
    pos_off[r] = tot_pos;
    tot_pos += L >= k + 1 ? L - k : 0;
  }
  pos_off[n_reads] = tot_pos;
  uint8_t *sol = xmalloc(tot_pos + 1);
  int64_t *item_off = xmalloc(sizeof(int64_t) * (n_reads + 1));
  uint32_t *items = NULL;
  int64_t n = 0;
  for (int pass = 0; pass < 3; ++pass) {
    if (pass == 2) {
      int64_t acc = 0;
      for (int64_t r = 0; r < n_reads; ++r) { int64_t c = item_off[r]; item_off[r] = acc; acc += c; }
      n = acc;
      items = xmalloc((size_t)n * W * 4 + 16);
    }
#pragma omp parallel num_threads(threads)
    {
      int64_t scap = 1024; uint8_t *s = xmalloc(scap), *rcs = xmalloc(scap);
      uint32_t a[64], b[64];
#pragma omp for schedule(dynamic, 2048)
      for (int64_t r = 0; r < n_reads; ++r) {
        int64_t L = starts[r + 1] - starts[r];
        if (L < k + 1) { if (pass == 1) item_off[r] = 0; continue; }
        if (L > scap) { scap = L * 2; s = xrealloc(s, scap); rcs = xrealloc(rcs, scap); }
        const uint8_t *c = bases + starts[r];
        for (int64_t j = 0; j < L; ++j) s[j] = c[L - 1 - j];
        for (int64_t j = 0; j < L; ++j) rcs[j] = 3 - s[L - 1 - j];
        int64_t npos = L - k;
        uint8_t *so = sol + pos_off[r];
        if (pass == 0) {
          for (int64_t i = 0; i < npos; ++i) {
            if (!solid) { so[i] = 1; continue; }
            pack_chars(s + i, k + 1, a, Wk);
            pack_chars(rcs + (L - 1 - i - k), k + 1, b, Wk);
            so[i] = (uint8_t)edge_is_solid(solid, cmp_words(b, a, Wk) < 0 ? b : a, Wk);
          }
          continue;
        }
        int64_t cnt = 0;
        uint32_t *o = pass == 2 ? items + item_off[r] * W : NULL;
#define PUSH_ITEM(ptr, nch, prevc)                                            \
  do {                                                                        \
    if (o) {                                                                  \
      pack_chars((ptr), (nch), o, W);                                         \
      o[W - 1] |= (uint32_t)((nch) == k) << 3;                                \
      o[W - 1] |= (uint32_t)(prevc);                                          \
      o += W;                                                                 \
    }                                                                         \
    ++cnt;                                                                    \
  } while (0)
        for (int64_t i = 0; i < npos; ++i) {
          if (!so[i]) continue;
          const uint8_t *e = s + i, *rc = rcs + (L - 1 - i - k);
          int pal = memcmp(e, rc, k + 1) == 0;
          int first = (i == 0) || !so[i - 1], lastp = (i == npos - 1) || !so[i + 1];
          PUSH_ITEM(e + 1, k, e[0]);
          if (!pal) PUSH_ITEM(rc + 1, k, rc[0]);
          if (first) {
            PUSH_ITEM(e, k, ORC_SENTINEL);
            if (!pal) PUSH_ITEM(rc + 2, k - 1, rc[1]);
          }
          if (lastp) {
            PUSH_ITEM(e + 2, k - 1, e[1]);
            if (!pal) PUSH_ITEM(rc, k, ORC_SENTINEL);
          }
        }
#undef PUSH_ITEM
        if (pass == 1) item_off[r] = cnt;
      }
      free(s); free(rcs);
    }
  }
  free(sol); free(pos_off); free(item_off);
  if (solid) orc_edges_free(solid);
  int64_t *bstart = xmalloc(sizeof(int64_t) * (ORC_NUM_BUCKETS + 1));
  sort_records(items, n, W, threads, bstart);
  orc_sdbg *g = sdbg_postprocess(items, n, &f, bstart, threads);
  free(bstart);
  free(items);
  return g;
}

/* SdbgWriter::Write / SdbgMeta::Serialize (src/sdbg/sdbg_writer.h, sdbg_meta.h):
 * item = uint16 (w | last<<4 | tip<<5 | min(mult,255)<<8) [+ uint16 mult if mult > 254] [+ tip label words if tip]. */
int orc_sdbg_write(const orc_sdbg *g, const char *prefix, int n_files) {
  if (n_files < 1) n_files = 1;
  char path[4096];
  FILE **fps = xmalloc(sizeof(FILE *) * n_files);
  int64_t *foff = xmalloc(sizeof(int64_t) * n_files);
  for (int f = 0; f < n_files; ++f) {
    snprintf(path, sizeof path, "%s.sdbg.%d", prefix, f);
    fps[f] = fopen(path, "wb");
    if (!fps[f]) return fail("cannot create %s", path);
    foff[f] = 0;
  }
  snprintf(path, sizeof path, "%s.sdbg_info", prefix);
  FILE *fi = fopen(path, "w");
  if (!fi) return fail("cannot create %s", path);
  fprintf(fi, "k %d\nwords_per_tip_label %d\nnum_buckets %d\nnum_files %d\n", g->k, g->words_per_tip, ORC_NUM_BUCKETS, n_files);
  /* SdbgMeta (src/sdbg/sdbg_meta.h): a bucket record nobody wrote to keeps bucket_id = kUninitializedBucketID = size_t(-1)
   * and zeros elsewhere; the records are sorted by bucket_id before they are serialised, so the unused ones come LAST. */
  int64_t pos = 0, tpos = 0, n_empty = 0;
  for (int b = 0; b < ORC_NUM_BUCKETS; ++b) {
    int f = (int)((int64_t)b * n_files / ORC_NUM_BUCKETS);
    int64_t c = g->bucket_items[b];
    if (!c) { ++n_empty; continue; }
    int64_t start = foff[f];
    for (int64_t i = pos; i < pos + c; ++i) {
      int m = g->mul[i];
      uint16_t rec = (uint16_t)(g->w[i] | (g->last[i] << 4) | (g->tip[i] << 5) | ((m > 255 ? 255 : m) << 8));
      fwrite(&rec, 2, 1, fps[f]); foff[f] += 2;
      if (m > 254) { uint16_t mm = (uint16_t)m; fwrite(&mm, 2, 1, fps[f]); foff[f] += 2; }
      if (g->tip[i]) { fwrite(g->tip_labels + tpos * g->words_per_tip, 4, g->words_per_tip, fps[f]); foff[f] += 4 * g->words_per_tip; ++tpos; }
    }
    fprintf(fi, "%d %d %lld %lld %lld %lld\n", b, f, (long long)start, (long long)c,
            (long long)g->bucket_tips[b], (long long)g->bucket_large[b]);
    pos += c;
  }
  for (int64_t i = 0; i < n_empty; ++i) fprintf(fi, "18446744073709551615 0 0 0 0 0\n");
  fprintf(fi, "item_count %lld\ntip_count %lld\nlarge_mul_count %lld\n", (long long)g->n, (long long)g->n_tips, (long long)g->n_large);
  fclose(fi);
  for (int f = 0; f < n_files; ++f) fclose(fps[f]);
  free(fps); free(foff);
  return 0;
}
orc_sdbg *orc_sdbg_read(const char *prefix) {
  char path[4096];
  snprintf(path, sizeof path, "%s.sdbg_info", prefix);
  FILE *fi = fopen(path, "r");
  if (!fi) { fail("cannot open %s", path); return NULL; }
  long long k, wpt, nb, nf;
  if (scan_field(fi, "k", &k) || scan_field(fi, "words_per_tip_label", &wpt) || scan_field(fi, "num_buckets", &nb) ||
      scan_field(fi, "num_files", &nf)) { fclose(fi); return NULL; }
  orc_sdbg *g = sdbg_new((int)k, 1024);
  uint8_t **fdata = xmalloc(sizeof(uint8_t *) * nf);
  for (int f = 0; f < nf; ++f) {
    snprintf(path, sizeof path, "%s.sdbg.%d", prefix, f);
    FILE *fp = fopen(path, "rb");
    if (!fp) { fail("cannot open %s", path); return NULL; }
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    fdata[f] = xmalloc(sz);
    if (fread(fdata[f], 1, sz, fp) != (size_t)sz) { fail("short read %s", path); return NULL; }
    fclose(fp);
  }
  for (int b = 0; b < nb; ++b) {
    unsigned long long bid;
    long long fid, off, items, tips, large;
    if (fscanf(fi, "%llu %lld %lld %lld %lld %lld", &bid, &fid, &off, &items, &tips, &large) != 6) { fail("bad bucket line"); return NULL; }
    if (bid >= (unsigned long long)ORC_NUM_BUCKETS || fid < 0 || items == 0) continue;   /* unused record */
    if (fid >= nf) { fail("bucket record names file %lld of %lld", fid, nf); return NULL; }
    const uint8_t *p = fdata[fid] + off;
    for (long long i = 0; i < items; ++i) {
      uint16_t rec; memcpy(&rec, p, 2); p += 2;
      int m = rec >> 8, tip = (rec >> 5) & 1;
      if (m == 255) { uint16_t mm; memcpy(&mm, p, 2); p += 2; m = mm; }
      uint32_t label[64];
      if (tip) { memcpy(label, p, 4 * wpt); p += 4 * wpt; }
      sdbg_push(g, (int)bid, rec & 15, (rec >> 4) & 1, tip, m, label);
    }
  }
  long long ic, tc, lc;
  if (scan_field(fi, "item_count", &ic) || scan_field(fi, "tip_count", &tc) || scan_field(fi, "large_mul_count", &lc)) { fclose(fi); return NULL; }
  fclose(fi);
  for (int f = 0; f < nf; ++f) free(fdata[f]);
  free(fdata);
  if (ic != g->n || tc != g->n_tips || lc != g->n_large) { fail("sdbg totals mismatch"); orc_sdbg_free(g); return NULL; }
  return g;
}
int orc_cmd_seq2sdbg(int k, int k_from, const char *input_prefix, const char *contig, const char *bubble,
                     const char *addi_contig, const char *local_contig, const char *out_prefix, int threads) {
  orc_seqs *s = orc_seqs_new();
  int rc = 0;
  if (input_prefix && *input_prefix) {
    orc_edges *e = orc_edges_read(input_prefix);
    if (!e) rc = -1; else { orc_seqs_add_edges(s, e); orc_edges_free(e); }
  }
  if (!rc && contig && *contig) rc = orc_seqs_add_contigs(s, contig, k + 1, 1, k_from, k);
  if (!rc && bubble && *bubble) rc = orc_seqs_add_contigs(s, bubble, k + 1, 1, k_from, k);
  if (!rc && addi_contig && *addi_contig) rc = orc_seqs_add_contigs(s, addi_contig, k + 1, 0, 0, 0);
  if (!rc && local_contig && *local_contig) rc = orc_seqs_add_contigs(s, local_contig, k + 1, 0, 0, 0);
  if (!rc) {
    orc_sdbg *g = orc_seq2sdbg(s, k, threads);
    if (!g) rc = -1; else { rc = orc_sdbg_write(g, out_prefix, threads); orc_sdbg_free(g); }
  }
  orc_seqs_free(s);
  return rc;
}
int orc_cmd_read2sdbg(const char *read_lib_file, int k, int min_count, const char *out_prefix, int threads) {
  orc_reads *r = load_lib(read_lib_file);
  if (!r) return -1;
  orc_sdbg *g = orc_read2sdbg(r->bases, r->starts, r->n, k, min_count, threads);
  orc_reads_free(r);
  if (!g) return -1;
  int rc = orc_sdbg_write(g, out_prefix, threads);
  orc_sdbg_free(g);
  return rc;
}
