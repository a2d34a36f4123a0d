// hostio.cu -- files in, files out, the device pipelines in between.  The on-disk layouts restate what
// megahit v1.2.9 reads and writes (EdgeWriter/EdgeIoMetadata, SdbgWriter/SdbgMeta, BinaryWriter, lib_info,
// ContigReader); the reference only fixes the file NAMES (assemble/assemble_wrapper.py:93-97,166-200,228-250).
#include "hostio.h"
#include <zlib.h>
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <thread>
#include "mfsdbg.h"

namespace mf {

namespace {

std::string errno_msg(const std::string &what, const std::string &path) {
  return what + " " + path + ": " + strerror(errno);
}
// whole BINARY file into memory: never through zlib, whose gzopen switches to gzip decoding on the magic bytes 1f 8b -- an edge
// word or a read length may well start with them
std::vector<uint8_t> slurp_binary(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) throw IoError(errno_msg("cannot open", path));
  std::vector<uint8_t> buf;
  size_t cap = 1 << 22, len = 0;
  buf.resize(cap);
  for (;;) {
    if (len == cap) { cap *= 2; buf.resize(cap); }
    const size_t got = fread(buf.data() + len, 1, cap - len, f);
    len += got;
    if (got == 0) {
      if (ferror(f)) { fclose(f); throw IoError("read error on " + path); }
      break;
    }
  }
  fclose(f);
  buf.resize(len);
  return buf;
}
// whole TEXT file (FASTA / FASTQ: plain, gzip or FIFO) into memory
std::vector<uint8_t> slurp(const std::string &path) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw IoError(errno_msg("cannot open", path));
  gzbuffer(f, 1 << 20);
  std::vector<uint8_t> buf;
  size_t cap = 1 << 22, len = 0;
  buf.resize(cap);
  for (;;) {
    if (len == cap) { cap *= 2; buf.resize(cap); }
    int want = (int)std::min<size_t>(cap - len, 1u << 30);
    int got = gzread(f, buf.data() + len, want);
    if (got < 0) { gzclose(f); throw IoError("read error on " + path); }
    if (got == 0) break;
    len += got;
  }
  gzclose(f);
  buf.resize(len);
  return buf;
}
struct File {
  FILE *fp = nullptr;
  std::string path;
  File(const std::string &p, const char *mode) : path(p) {
    fp = fopen(p.c_str(), mode);
    if (!fp) throw IoError(errno_msg(mode[0] == 'r' ? "cannot open" : "cannot create", p));
  }
  ~File() { if (fp) fclose(fp); }
  void write(const void *d, size_t n) {
    if (n && fwrite(d, 1, n, fp) != n) throw IoError(errno_msg("write failed on", path));
  }
  void close() {
    if (fp && fclose(fp) != 0) { fp = nullptr; throw IoError(errno_msg("close failed on", path)); }
    fp = nullptr;
  }
};
struct DevMem {   // scoped cudaMalloc
  void *p = nullptr;
  explicit DevMem(size_t bytes) { MF_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16))); }
  ~DevMem() { if (p) cudaFree(p); }
  template <class T> T *as() { return reinterpret_cast<T *>(p); }
};
int bucket_file(int bucket, int n_files) { return (int)((int64_t)bucket * n_files / kNumBuckets); }

}  // namespace

// ---------------------------------------------------------------- buildlib
namespace {
// text bytes per file and chunk (pinned, two sets: one is read while one is packed); MFSDBG_TEXT_CHUNK overrides (tests)
static size_t text_chunk() {
  const char *v = getenv("MFSDBG_TEXT_CHUNK");
  return v && *v ? (size_t)std::max(1024ll, atoll(v)) : (size_t)192 << 20;
}

// A FASTQ / FASTA text file (plain, gzip or FIFO) read in chunks that end on record boundaries (single-line records: 4 lines
// per FASTQ record, 2 per FASTA record -- what the device parser accepts).
struct TextStream {
  gzFile f = nullptr;
  std::string path;
  std::vector<uint8_t> carry;   // the partial record (or the records handed back) behind the last chunk
  bool eof = false;
  int lines_per_rec = 0;
  explicit TextStream(const std::string &p) : path(p) {
    f = gzopen(p.c_str(), "rb");
    if (!f) throw IoError(errno_msg("cannot open", p));
    gzbuffer(f, 1 << 20);
  }
  ~TextStream() { if (f) gzclose(f); }
  // fills buf[0 .. cap) with whole records; returns the bytes used and their record count; what is left over stays in `carry`
  size_t next(uint8_t *buf, size_t cap, int64_t *n_rec) {
    size_t len = carry.size();
    if (len > cap) throw IoError("a single record of " + path + " exceeds the chunk size");
    if (len) memcpy(buf, carry.data(), len);
    carry.clear();
    while (!eof && len < cap) {
      const int got = gzread(f, buf + len, (unsigned)std::min<size_t>(cap - len, 1u << 30));
      if (got < 0) throw IoError("read error on " + path);
      if (got == 0) eof = true;
      len += (size_t)got;
    }
    if (lines_per_rec == 0 && len > 0) lines_per_rec = buf[0] == '>' ? 2 : 4;
    if (eof && len > 0 && buf[len - 1] != '\n' && len < cap) buf[len++] = '\n';   // last line without its newline
    // newlines of the chunk; the cut goes behind the last one that completes a record
    int64_t lines = 0;
    size_t last_rec_end = 0;
    for (const uint8_t *q = buf, *end = buf + len; q < end;) {
      const uint8_t *nl = (const uint8_t *)memchr(q, '\n', (size_t)(end - q));
      if (!nl) break;
      ++lines;
      if (lines_per_rec && lines % lines_per_rec == 0) last_rec_end = (size_t)(nl - buf) + 1;
      q = nl + 1;
    }
    if (last_rec_end < len) {
      if (eof) throw IoError("read file is not made of whole single-line FASTQ/FASTA records (multi-line records are not supported): " + path);
      carry.assign(buf + last_rec_end, buf + len);
    }
    *n_rec = lines_per_rec ? (lines - lines % lines_per_rec) / lines_per_rec : 0;
    if (last_rec_end == 0 && !eof && len == cap) throw IoError("a single record of " + path + " exceeds the chunk size");
    return last_rec_end;
  }
  // keep only the first `keep` records of buf[0 .. len): the rest goes back in front of the carry.  Returns the bytes kept.
  size_t give_back(const uint8_t *buf, size_t len, int64_t keep) {
    int64_t lines = 0;
    size_t cut = 0;
    const int64_t want = keep * lines_per_rec;
    for (const uint8_t *q = buf, *end = buf + len; q < end && lines < want;) {
      const uint8_t *nl = (const uint8_t *)memchr(q, '\n', (size_t)(end - q));
      if (!nl) break;
      ++lines;
      cut = (size_t)(nl - buf) + 1;
      q = nl + 1;
    }
    if (want == 0) cut = 0;
    std::vector<uint8_t> rest(buf + cut, buf + len);
    rest.insert(rest.end(), carry.begin(), carry.end());
    carry.swap(rest);
    return cut;
  }
  bool done() const { return eof && carry.empty(); }
};
struct ChunkSet {   // one chunk of every file of a library, in pinned memory
  HostBuf buf[2];
  size_t len[2] = {0, 0};
  int64_t n_rec = 0;
  bool last = false;
};
}  // namespace

// Streams the library through the GPU packer: while chunk i is copied, packed and written, a helper thread reads chunk i+1
// (zlib / disk) into the other pinned set -- neither the text nor the packed reads are ever held whole (round 1 slurped every
// file into host memory and then into HBM: 11 GB of text twice for the 5 Gbp sample).
void file_buildlib(Ctx &c, const char *lib_file, const char *out_prefix, int n_policy) {
  std::ifstream lib(lib_file);
  if (!lib) throw IoError(errno_msg("cannot open", lib_file));
  File bin(std::string(out_prefix) + ".bin", "wb");
  std::ostringstream info;
  int64_t total_reads = 0, total_bases = 0;
  std::string meta, spec;
  ChunkSet sets[2];
  DevBuf d_text[2];
  try {
    while (std::getline(lib, meta)) {
      if (!std::getline(lib, spec)) break;
      std::istringstream ss(spec);
      std::string type, f1, f2;
      ss >> type >> f1 >> f2;
      std::vector<std::unique_ptr<TextStream>> in;
      bool paired = false;
      if (type == "pe" && !f2.empty()) {
        // open both before reading either: MitoFlex may hand over two FIFOs fed by `gzip -dc` children
        in.emplace_back(new TextStream(f1));
        in.emplace_back(new TextStream(f2));
        paired = true;
      } else if ((type == "se" || type == "interleaved") && !f1.empty()) {
        in.emplace_back(new TextStream(f1));
        paired = type == "interleaved";
      } else {
        throw IoError("bad library line in " + std::string(lib_file) + ": " + spec);
      }
      const int nf = (int)in.size();
      const size_t kTextChunk = text_chunk();
      for (auto &st : sets)
        for (int f = 0; f < nf; ++f) st.buf[f].reserve(kTextChunk + 16);
      auto read_chunk = [&](ChunkSet *st) {
        int64_t nr[2] = {0, 0};
        for (int f = 0; f < nf; ++f) st->len[f] = in[f]->next(st->buf[f].as<uint8_t>(), kTextChunk, &nr[f]);
        if (nf == 2 && nr[0] != nr[1]) {   // the files' chunks hold different numbers of records: both keep the smaller count
          const int big = nr[0] > nr[1] ? 0 : 1;
          st->len[big] = in[big]->give_back(st->buf[big].as<uint8_t>(), st->len[big], nr[1 - big]);
          nr[big] = nr[1 - big];
        }
        st->n_rec = nr[0];
        st->last = true;
        for (int f = 0; f < nf; ++f) st->last = st->last && in[f]->done();
        if (st->last && nf == 2 && (in[0]->done() != in[1]->done())) throw IoError("paired files have different numbers of reads");
        if (st->n_rec == 0 && !st->last) throw IoError("paired files have different numbers of reads");
      };
      const int64_t first_read = total_reads;
      int lib_max_len = 0, cur = 0;
      read_chunk(&sets[cur]);
      for (;;) {
        ChunkSet &st = sets[cur];
        std::exception_ptr rerr;
        std::thread reader;
        if (!st.last) reader = std::thread([&, cur] { try { read_chunk(&sets[cur ^ 1]); } catch (...) { rerr = std::current_exception(); } });
        try {
          if (st.n_rec > 0) {
            const uint8_t *ptrs[2] = {nullptr, nullptr};
            int64_t sizes[2] = {0, 0};
            for (int f = 0; f < nf; ++f) {
              d_text[f].reserve(st.len[f] + 64);
              MF_CUDA(cudaMemcpyAsync(d_text[f].p, st.buf[f].p, st.len[f], cudaMemcpyHostToDevice, c.stream));
              ptrs[f] = d_text[f].as<uint8_t>();
              sizes[f] = (int64_t)st.len[f];
            }
            ReadsView r;
            int max_len = 0;
            dev_pack_fastq(c, ptrs, sizes, nf, n_policy, &r, &max_len);
            std::vector<uint32_t> stream;
            reads_to_bin_stream(c, r, &stream);
            bin.write(stream.data(), stream.size() * 4);
            lib_max_len = std::max(lib_max_len, max_len);
            total_reads += r.n_reads;
            total_bases += r.n_bases;
          }
        } catch (...) {
          if (reader.joinable()) reader.join();
          throw;
        }
        if (reader.joinable()) reader.join();
        if (rerr) std::rethrow_exception(rerr);
        if (st.last) break;
        cur ^= 1;
      }
      info << meta << '\n' << (paired ? "pe " : "se ") << first_read << ' ' << total_reads - 1 << ' ' << lib_max_len << '\n';
    }
  } catch (...) {
    for (auto &d : d_text) d.release();
    for (auto &st : sets)
      for (auto &b : st.buf) b.release();
    throw;
  }
  for (auto &d : d_text) d.release();
  for (auto &st : sets)
    for (auto &b : st.buf) b.release();
  bin.close();
  File fi(std::string(out_prefix) + ".lib_info", "w");   // meta file last
  std::string head = std::to_string(total_bases) + " " + std::to_string(total_reads) + "\n" + info.str();
  fi.write(head.data(), head.size());
  fi.close();
}

static void load_read_lib(Ctx &c, const char *read_lib_file, ReadsView *r) {
  std::vector<uint8_t> raw = slurp_binary(std::string(read_lib_file) + ".bin");
  if (raw.size() % 4) throw IoError("read library .bin is not a whole number of words");
  bin_stream_to_reads(c, reinterpret_cast<const uint32_t *>(raw.data()), (int64_t)(raw.size() / 4), r);
}

// ---------------------------------------------------------------- edges files
static void write_edges(Ctx &c, const EdgesView &e, const std::string &prefix, int n_files) {
  std::vector<uint32_t> host((size_t)e.n_edges * e.words);
  if (e.n_edges) c.d2h(host.data(), e.edges, host.size() * 4);
  std::vector<std::unique_ptr<File>> files;
  for (int f = 0; f < n_files; ++f) files.emplace_back(new File(prefix + ".edges." + std::to_string(f), "wb"));
  std::ostringstream info;
  info << "kmer_size " << e.k << "\nwords_per_edge " << e.words << "\nnum_files " << n_files << "\nnum_buckets " << kNumBuckets
       << "\nnum_edges " << e.n_edges << "\nis_sorted 1\n";
  std::vector<int64_t> foff(n_files, 0);
  int64_t pos = 0;
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t cnt = c.edge_bucket_counts[b];
    if (!cnt) { info << b << " -1 0 0\n"; continue; }
    const int f = bucket_file(b, n_files);
    files[f]->write(host.data() + pos * e.words, (size_t)cnt * e.words * 4);
    info << b << ' ' << f << ' ' << foff[f] << ' ' << cnt << '\n';
    foff[f] += cnt;
    pos += cnt;
  }
  if (pos != e.n_edges) throw std::runtime_error("bucket counts do not add up to the edge count");
  for (auto &f : files) f->close();
  File fi(prefix + ".edges.info", "w");   // meta file last: a failed run leaves nothing parseable
  const std::string s = info.str();
  fi.write(s.data(), s.size());
  fi.close();
}
struct HostEdges {
  int k = 0, words = 0;
  bool sorted = false;
  int64_t n = 0;
  std::vector<uint32_t> data;
};
static void expect_field(std::istream &is, const char *name, long long *v) {
  std::string tok;
  if (!(is >> tok >> *v) || tok != name) throw IoError(std::string("malformed edges/sdbg info: expected ") + name);
}
static HostEdges read_edges(const std::string &prefix) {
  std::ifstream is(prefix + ".edges.info");
  if (!is) throw IoError(errno_msg("cannot open", prefix + ".edges.info"));
  long long k, words, nfiles, nbuckets, nedges, sorted;
  expect_field(is, "kmer_size", &k);
  expect_field(is, "words_per_edge", &words);
  expect_field(is, "num_files", &nfiles);
  expect_field(is, "num_buckets", &nbuckets);
  expect_field(is, "num_edges", &nedges);
  expect_field(is, "is_sorted", &sorted);
  // a mismatched or corrupt meta file must end in a clean EIO, not in a wrong stride on the device or a huge allocation
  if (k < 1 || k > 255 || words != words_edge((int)k)) throw IoError("edges.info: words_per_edge does not belong to kmer_size");
  if (sorted ? nbuckets != kNumBuckets : (nbuckets < 0 || nbuckets > kNumBuckets))   // unsorted (iterate) files carry no bucket table
    throw IoError("edges.info: num_buckets must be 65536 for sorted edges");
  if (nfiles < 1 || nfiles > 65536 || nedges < 0 || nedges > ((long long)1 << 40)) throw IoError("edges.info: implausible num_files / num_edges");
  HostEdges e;
  e.k = (int)k; e.words = (int)words; e.sorted = sorted != 0; e.n = nedges;
  e.data.resize((size_t)nedges * words);
  std::vector<std::vector<uint8_t>> fdata(nfiles);
  for (int f = 0; f < nfiles; ++f) fdata[f] = slurp_binary(prefix + ".edges." + std::to_string(f));
  int64_t pos = 0;
  const size_t rec = (size_t)words * 4;
  if (e.sorted) {
    for (int b = 0; b < nbuckets; ++b) {
      long long bid, fid, off, cnt;
      if (!(is >> bid >> fid >> off >> cnt) || bid != b) throw IoError("Invalid format: bucket id not matched!");
      if (fid < 0 || cnt == 0) continue;
      if (off < 0 || cnt < 0) throw IoError("edges.info: negative offset or count");
      if (fid >= nfiles || (size_t)(off + cnt) * rec > fdata[fid].size() || pos + cnt > nedges) throw IoError("edge file shorter than its meta says");
      memcpy(e.data.data() + pos * words, fdata[fid].data() + (size_t)off * rec, (size_t)cnt * rec);
      pos += cnt;
    }
  } else {
    for (int f = 0; f < nfiles; ++f) {
      const int64_t cnt = (int64_t)(fdata[f].size() / rec);
      if (pos + cnt > nedges) throw IoError("more edges on disk than the meta says");
      memcpy(e.data.data() + pos * words, fdata[f].data(), (size_t)cnt * rec);
      pos += cnt;
    }
  }
  if (pos != nedges) throw IoError("edge count mismatch between meta and files");
  return e;
}

void file_count(Ctx &c, const char *read_lib_file, int k, int min_count, const char *out_prefix, int n_files) {
  ReadsView r;
  load_read_lib(c, read_lib_file, &r);
  EdgesView e;
  std::vector<int64_t> counting(kNumBuckets, 0);
  c.begin_call();
  dev_count(c, r, k, min_count, &e, counting.data());
  c.end_call();
  write_edges(c, e, out_prefix, n_files);
  // KmerCounter::Lv0Postprocess: cumulative distinct-edge histogram
  File fc(std::string(out_prefix) + ".counting", "w");
  std::ostringstream ss;
  long long acc = 0;
  for (int i = 1; i <= kMaxMul; ++i) { acc += counting[i]; ss << i << ' ' << acc << '\n'; }
  const std::string s = ss.str();
  fc.write(s.data(), s.size());
  fc.close();
}

// ---------------------------------------------------------------- contigs
namespace {
struct HostSeqs {
  std::vector<uint32_t> packed;    // 2-bit, back to back, stored (reversed) orientation
  std::vector<int64_t> starts{0};
  std::vector<uint16_t> mult;
  std::vector<int64_t> item_base{0};
  void add(const std::vector<uint8_t> &codes, int m, int k) {
    const int64_t s = starts.back();
    const int64_t L = (int64_t)codes.size();
    packed.resize((size_t)((s + L + 15) >> 4) + 1, 0u);
    for (int64_t i = 0; i < L; ++i) {
      const int64_t g = s + i;
      packed[g >> 4] |= (uint32_t)codes[i] << (30 - 2 * (g & 15));
    }
    starts.push_back(s + L);
    mult.push_back((uint16_t)m);
    item_base.push_back(item_base.back() + (L >= k + 1 ? 2 * (L - k + 2) : 0));
  }
};
inline uint8_t code_of(char ch) {
  switch (ch) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': case 'N': case 'n': return 2;
    case 'T': case 't': return 3;
    default: return 0;
  }
}
// ContigReader: ">id flag=F multi=M len=L", sequence on the following line(s); contig_reverse = true;
// loop contigs (flag & 2) get their first k_to-k_from bases appended; multiplicity = min(65535, int(M + 0.5)).
void add_contigs(HostSeqs *hs, const std::string &path, int k, bool extend_loop, int k_from, int k_to) {
  std::vector<uint8_t> raw = slurp(path);
  const int min_len = k + 1;
  size_t p = 0, n = raw.size();
  std::vector<uint8_t> codes;
  while (p < n) {
    while (p < n && raw[p] != '>' && raw[p] != '@') { while (p < n && raw[p] != '\n') ++p; if (p < n) ++p; }
    if (p >= n) break;
    const bool fq = raw[p] == '@';
    size_t h = p + 1;
    while (p < n && raw[p] != '\n') ++p;
    std::string header((const char *)raw.data() + h, p - h);
    if (p < n) ++p;
    codes.clear();
    while (p < n && raw[p] != '>' && raw[p] != '+' && raw[p] != '@') {
      while (p < n && raw[p] != '\n') { if (raw[p] != '\r') codes.push_back(code_of((char)raw[p])); ++p; }
      if (p < n) ++p;
    }
    if (fq && p < n && raw[p] == '+') {   // skip quality
      while (p < n && raw[p] != '\n') ++p;
      if (p < n) ++p;
      size_t q = 0;
      while (p < n && q < codes.size()) { while (p < n && raw[p] != '\n') { if (raw[p] != '\r') ++q; ++p; } if (p < n) ++p; }
    }
    unsigned flag = 0;
    float multi = 1.0f;
    size_t sp = header.find_first_of(" \t");
    if (sp != std::string::npos) {
      std::string comment = header.substr(header.find_first_not_of(" \t", sp) == std::string::npos ? header.size() : header.find_first_not_of(" \t", sp));
      sscanf(comment.c_str(), "flag=%u multi=%f", &flag, &multi);
    }
    size_t ext = (extend_loop && (flag & 2u)) ? (size_t)std::max(0, k_to - k_from) : 0;
    ext = std::min(ext, codes.size());
    if ((int64_t)(codes.size() + ext) < min_len) continue;
    for (size_t i = 0; i < ext; ++i) codes.push_back(codes[i]);
    std::reverse(codes.begin(), codes.end());
    int m = (int)(multi + 0.5f);
    hs->add(codes, std::min(m, kMaxMul), k);
  }
}
}  // namespace

// ---------------------------------------------------------------- sdbg files
// SdbgWriter::Write: uint16 (w | last<<4 | tip<<5 | min(mult,255)<<8) [+ uint16 mult if > 254] [+ tip label words]
static void write_sdbg(Ctx &c, const SdbgView &g, const std::string &prefix, int n_files) {
  std::vector<uint32_t> rec((size_t)g.n_items), labels((size_t)g.n_tips * g.words_tip);
  if (g.n_items) c.d2h(rec.data(), g.rec, rec.size() * 4);
  if (g.n_tips) c.d2h(labels.data(), g.labels, labels.size() * 4);
  std::vector<std::unique_ptr<File>> files;
  for (int f = 0; f < n_files; ++f) files.emplace_back(new File(prefix + ".sdbg." + std::to_string(f), "wb"));
  std::ostringstream info;
  info << "k " << g.k << "\nwords_per_tip_label " << g.words_tip << "\nnum_buckets " << kNumBuckets << "\nnum_files " << n_files << '\n';
  std::vector<int64_t> foff(n_files, 0);
  std::vector<uint8_t> buf;
  // SdbgMeta: a bucket record nobody wrote to keeps bucket_id = kUninitializedBucketID = size_t(-1) and zeros elsewhere, and
  // the records are sorted by bucket_id before they are serialised -- the unused ones come LAST (recollection shared by the
  // oracle; the first run against a real megahit_core decides it)
  int64_t pos = 0, tpos = 0, large = 0, n_empty = 0;
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t items = c.sdbg_bucket_stats[(size_t)b * 3], tips = c.sdbg_bucket_stats[(size_t)b * 3 + 1],
                  lg = c.sdbg_bucket_stats[(size_t)b * 3 + 2];
    if (!items) { ++n_empty; continue; }
    const int f = bucket_file(b, n_files);
    buf.clear();
    for (int64_t i = pos; i < pos + items; ++i) {
      const uint32_t r = rec[(size_t)i];
      const uint32_t m = r >> 8;
      const uint16_t small = (uint16_t)((r & 0x3f) | (std::min<uint32_t>(m, 255) << 8));
      buf.insert(buf.end(), (const uint8_t *)&small, (const uint8_t *)&small + 2);
      if (m > 254) { const uint16_t mm = (uint16_t)m; buf.insert(buf.end(), (const uint8_t *)&mm, (const uint8_t *)&mm + 2); ++large; }
      if (r & 0x20) {
        const uint8_t *lp = (const uint8_t *)(labels.data() + (size_t)tpos * g.words_tip);
        buf.insert(buf.end(), lp, lp + 4 * g.words_tip);
        ++tpos;
      }
    }
    files[f]->write(buf.data(), buf.size());
    info << b << ' ' << f << ' ' << foff[f] << ' ' << items << ' ' << tips << ' ' << lg << '\n';
    foff[f] += (int64_t)buf.size();
    pos += items;
  }
  if (pos != g.n_items || tpos != g.n_tips) throw std::runtime_error("sdbg bucket statistics do not add up");
  for (int64_t i = 0; i < n_empty; ++i) info << "18446744073709551615 0 0 0 0 0\n";
  info << "item_count " << g.n_items << "\ntip_count " << g.n_tips << "\nlarge_mul_count " << large << '\n';
  for (auto &f : files) f->close();
  File fi(prefix + ".sdbg_info", "w");
  const std::string s = info.str();
  fi.write(s.data(), s.size());
  fi.close();
}

void file_seq2sdbg(Ctx &c, int k, int k_from, const char *input_prefix, const char *contig, const char *bubble,
                   const char *addi_contig, const char *local_contig, const char *out_prefix, int n_files) {
  HostEdges he;
  if (input_prefix && *input_prefix) {
    he = read_edges(input_prefix);
    if (he.k != k) throw IoError("edges were built for k=" + std::to_string(he.k) + ", not " + std::to_string(k));
  }
  HostSeqs hs;
  if (contig && *contig) add_contigs(&hs, contig, k, true, k_from, k);
  if (bubble && *bubble) add_contigs(&hs, bubble, k, true, k_from, k);
  if (addi_contig && *addi_contig) add_contigs(&hs, addi_contig, k, false, 0, 0);
  if (local_contig && *local_contig) add_contigs(&hs, local_contig, k, false, 0, 0);
  const int nseq = (int)hs.mult.size();
  DevMem d_edges(he.data.size() * 4 + 64), d_packed(hs.packed.size() * 4 + 256), d_starts(sizeof(int64_t) * (nseq + 1)),
      d_mult(sizeof(uint16_t) * (nseq + 1)), d_ibase(sizeof(int64_t) * (nseq + 1));
  if (he.n) c.h2d(d_edges.p, he.data.data(), he.data.size() * 4);
  SeqsView sv;
  if (nseq) {
    MF_CUDA(cudaMemsetAsync(d_packed.p, 0, hs.packed.size() * 4 + 256, c.stream));
    c.h2d(d_packed.p, hs.packed.data(), hs.packed.size() * 4);
    c.h2d(d_starts.p, hs.starts.data(), sizeof(int64_t) * (nseq + 1));
    c.h2d(d_mult.p, hs.mult.data(), sizeof(uint16_t) * nseq);
    c.h2d(d_ibase.p, hs.item_base.data(), sizeof(int64_t) * (nseq + 1));
    sv.packed = d_packed.as<uint32_t>();
    sv.starts = d_starts.as<int64_t>();
    sv.mult = d_mult.as<uint16_t>();
    sv.item_base = d_ibase.as<int64_t>();
    sv.nseq = nseq;
    sv.n_items = hs.item_base.back();
  }
  SdbgView g;
  c.begin_call();
  dev_seq2sdbg(c, d_edges.as<uint32_t>(), he.n, sv, k, 0, &g);
  c.end_call();
  write_sdbg(c, g, out_prefix, n_files);
}

void file_read2sdbg(Ctx &c, const char *read_lib_file, int k, int min_count, const char *out_prefix, int n_files) {
  ReadsView r;
  load_read_lib(c, read_lib_file, &r);
  EdgesView e;
  SdbgView g;
  c.begin_call();
  dev_count(c, r, k, min_count, &e, nullptr);
  dev_seq2sdbg(c, e.edges, e.n_edges, SeqsView{}, k, 1, &g);
  c.end_call();
  write_sdbg(c, g, out_prefix, n_files);
}


// ---------------------------------------------------------------- several GPUs (multi.cu): one file per GPU
// Rank r holds a contiguous piece of the sorted stream; its buckets go to file r in order, the meta rows are the same as in the
// single-GPU writers (bucket -> file, offset, count).
static void write_edges_multi(MultiGpu &mg, const std::string &prefix) {
  const int G = mg.world();
  int64_t total = 0;
  for (int r = 0; r < G; ++r) total += mg.edges(r).n_edges;
  const EdgesView &e0 = mg.edges(0);
  std::ostringstream info;
  info << "kmer_size " << e0.k << "\nwords_per_edge " << e0.words << "\nnum_files " << G << "\nnum_buckets " << kNumBuckets
       << "\nnum_edges " << total << "\nis_sorted 1\n";
  std::vector<std::string> rows(kNumBuckets);
  for (int r = 0; r < G; ++r) {
    Ctx &c = mg.ctx(r);
    const EdgesView &e = mg.edges(r);
    MF_CUDA(cudaSetDevice(c.device));
    std::vector<uint32_t> host((size_t)e.n_edges * e.words);
    if (e.n_edges) c.d2h(host.data(), e.edges, host.size() * 4);
    File f(prefix + ".edges." + std::to_string(r), "wb");
    f.write(host.data(), host.size() * 4);
    f.close();
    int64_t off = 0;
    for (int b = 0; b < kNumBuckets; ++b) {
      const int64_t cnt = c.edge_bucket_counts[b];
      if (!cnt) continue;
      if (!rows[b].empty()) throw std::runtime_error("bucket " + std::to_string(b) + " is held by two GPUs");
      rows[b] = std::to_string(b) + ' ' + std::to_string(r) + ' ' + std::to_string(off) + ' ' + std::to_string(cnt) + '\n';
      off += cnt;
    }
    if (off != e.n_edges) throw std::runtime_error("bucket counts do not add up to the edge count");
  }
  for (int b = 0; b < kNumBuckets; ++b) {
    if (rows[b].empty()) info << b << " -1 0 0\n";
    else info << rows[b];
  }
  File fi(prefix + ".edges.info", "w");   // meta file last
  const std::string s = info.str();
  fi.write(s.data(), s.size());
  fi.close();
}
static void write_sdbg_multi(MultiGpu &mg, const std::string &prefix) {
  const int G = mg.world();
  const SdbgView &g0 = mg.sdbg(0);
  std::ostringstream info;
  info << "k " << g0.k << "\nwords_per_tip_label " << g0.words_tip << "\nnum_buckets " << kNumBuckets << "\nnum_files " << G << '\n';
  std::vector<std::string> rows(kNumBuckets);
  int64_t n_items = 0, n_tips = 0, n_large = 0;
  std::vector<uint8_t> buf;
  for (int r = 0; r < G; ++r) {
    Ctx &c = mg.ctx(r);
    const SdbgView &g = mg.sdbg(r);
    MF_CUDA(cudaSetDevice(c.device));
    std::vector<uint32_t> rec((size_t)g.n_items), labels((size_t)g.n_tips * g.words_tip);
    if (g.n_items) c.d2h(rec.data(), g.rec, rec.size() * 4);
    if (g.n_tips) c.d2h(labels.data(), g.labels, labels.size() * 4);
    File f(prefix + ".sdbg." + std::to_string(r), "wb");
    int64_t pos = 0, tpos = 0, foff = 0;
    for (int b = 0; b < kNumBuckets; ++b) {
      const int64_t items = c.sdbg_bucket_stats[(size_t)b * 3], tips = c.sdbg_bucket_stats[(size_t)b * 3 + 1],
                    lg = c.sdbg_bucket_stats[(size_t)b * 3 + 2];
      if (!items) continue;
      if (!rows[b].empty()) throw std::runtime_error("sdbg bucket " + std::to_string(b) + " is held by two GPUs");
      buf.clear();
      for (int64_t i = pos; i < pos + items; ++i) {
        const uint32_t rr = rec[(size_t)i];
        const uint32_t m = rr >> 8;
        const uint16_t small = (uint16_t)((rr & 0x3f) | (std::min<uint32_t>(m, 255) << 8));
        buf.insert(buf.end(), (const uint8_t *)&small, (const uint8_t *)&small + 2);
        if (m > 254) { const uint16_t mm = (uint16_t)m; buf.insert(buf.end(), (const uint8_t *)&mm, (const uint8_t *)&mm + 2); ++n_large; }
        if (rr & 0x20) {
          const uint8_t *lp = (const uint8_t *)(labels.data() + (size_t)tpos * g.words_tip);
          buf.insert(buf.end(), lp, lp + 4 * g.words_tip);
          ++tpos;
        }
      }
      f.write(buf.data(), buf.size());
      rows[b] = std::to_string(b) + ' ' + std::to_string(r) + ' ' + std::to_string(foff) + ' ' + std::to_string(items) + ' ' +
                std::to_string(tips) + ' ' + std::to_string(lg) + '\n';
      foff += (int64_t)buf.size();
      pos += items;
    }
    if (pos != g.n_items || tpos != g.n_tips) throw std::runtime_error("sdbg bucket statistics do not add up");
    f.close();
    n_items += g.n_items;
    n_tips += g.n_tips;
  }
  int64_t n_empty = 0;
  for (int b = 0; b < kNumBuckets; ++b) {
    if (rows[b].empty()) ++n_empty;
    else info << rows[b];
  }
  for (int64_t i = 0; i < n_empty; ++i) info << "18446744073709551615 0 0 0 0 0\n";
  info << "item_count " << n_items << "\ntip_count " << n_tips << "\nlarge_mul_count " << n_large << '\n';
  File fi(prefix + ".sdbg_info", "w");
  const std::string s = info.str();
  fi.write(s.data(), s.size());
  fi.close();
}
// the read library cut into `world` runs of whole reads (by read count), each unpacked on its GPU
static std::vector<ReadsView> load_read_lib_multi(MultiGpu &mg, const char *read_lib_file) {
  std::vector<uint8_t> raw = slurp_binary(std::string(read_lib_file) + ".bin");
  if (raw.size() % 4) throw IoError("read library .bin is not a whole number of words");
  const uint32_t *st = reinterpret_cast<const uint32_t *>(raw.data());
  const int64_t nw = (int64_t)(raw.size() / 4);
  std::vector<int64_t> rec_off;
  for (int64_t p = 0; p < nw;) {
    rec_off.push_back(p);
    p += 1 + (((int64_t)st[p] + 15) >> 4);
    if (p > nw) throw IoError("truncated read library (.bin)");
  }
  const int64_t n = (int64_t)rec_off.size();
  rec_off.push_back(nw);
  const int G = mg.world();
  std::vector<ReadsView> out(G);
  for (int r = 0; r < G; ++r) {
    const int64_t lo = n * r / G, hi = n * (r + 1) / G;
    MF_CUDA(cudaSetDevice(mg.ctx(r).device));
    bin_stream_to_reads(mg.ctx(r), st + rec_off[lo], rec_off[hi] - rec_off[lo], &out[r]);
  }
  return out;
}
void file_count_multi(const std::vector<int> &devices, const char *read_lib_file, int k, int min_count, const char *out_prefix) {
  MultiGpu mg(devices);
  std::vector<ReadsView> reads = load_read_lib_multi(mg, read_lib_file);
  mg.count(reads, k, min_count, true);
  write_edges_multi(mg, out_prefix);
  File fc(std::string(out_prefix) + ".counting", "w");
  std::ostringstream ss;
  long long acc = 0;
  for (int i = 1; i <= kMaxMul; ++i) {
    for (int r = 0; r < mg.world(); ++r) acc += mg.counting(r)[i];   // the GPUs hold disjoint key ranges: the histograms add
    ss << i << ' ' << acc << '\n';
  }
  const std::string s = ss.str();
  fc.write(s.data(), s.size());
  fc.close();
}
void file_read2sdbg_multi(const std::vector<int> &devices, const char *read_lib_file, int k, int min_count, const char *out_prefix) {
  MultiGpu mg(devices);
  std::vector<ReadsView> reads = load_read_lib_multi(mg, read_lib_file);
  mg.read2sdbg(reads, k, min_count);
  write_sdbg_multi(mg, out_prefix);
}
void file_seq2sdbg_multi(const std::vector<int> &devices, int k, int k_from, const char *input_prefix, const char *contig, const char *bubble,
                         const char *addi_contig, const char *local_contig, const char *out_prefix) {
  HostEdges he;
  if (input_prefix && *input_prefix) {
    he = read_edges(input_prefix);
    if (he.k != k) throw IoError("edges were built for k=" + std::to_string(he.k) + ", not " + std::to_string(k));
  }
  HostSeqs all;
  if (contig && *contig) add_contigs(&all, contig, k, true, k_from, k);
  if (bubble && *bubble) add_contigs(&all, bubble, k, true, k_from, k);
  if (addi_contig && *addi_contig) add_contigs(&all, addi_contig, k, false, 0, 0);
  if (local_contig && *local_contig) add_contigs(&all, local_contig, k, false, 0, 0);
  MultiGpu mg(devices);
  const int G = mg.world(), nseq = (int)all.mult.size(), words = std::max(he.words, 1);
  std::vector<std::unique_ptr<DevMem>> keep;
  std::vector<const uint32_t *> d_edges(G, nullptr);
  std::vector<int64_t> n_edges(G, 0);
  std::vector<SeqsView> seqs(G);
  for (int r = 0; r < G; ++r) {
    Ctx &c = mg.ctx(r);
    MF_CUDA(cudaSetDevice(c.device));
    // edges by index, sequences by index: any split works, the exchange routes every item to the GPU that owns its prefix
    const int64_t e_lo = he.n * r / G, e_hi = he.n * (r + 1) / G;
    n_edges[r] = e_hi - e_lo;
    keep.emplace_back(new DevMem((size_t)n_edges[r] * words * 4 + 64));
    if (n_edges[r]) c.h2d(keep.back()->p, he.data.data() + (size_t)e_lo * words, (size_t)n_edges[r] * words * 4);
    d_edges[r] = keep.back()->as<uint32_t>();
    const int s_lo = (int)((int64_t)nseq * r / G), s_hi = (int)((int64_t)nseq * (r + 1) / G);
    if (s_hi > s_lo) {
      HostSeqs part;   // re-packed so that the piece starts at base 0
      for (int i = s_lo; i < s_hi; ++i) {
        std::vector<uint8_t> codes((size_t)(all.starts[i + 1] - all.starts[i]));
        for (size_t j = 0; j < codes.size(); ++j) {
          const int64_t g = all.starts[i] + (int64_t)j;
          codes[j] = (uint8_t)((all.packed[g >> 4] >> (30 - 2 * (g & 15))) & 3u);
        }
        part.add(codes, all.mult[i], k);
      }
      const int ns = s_hi - s_lo;
      keep.emplace_back(new DevMem(part.packed.size() * 4 + 256));
      MF_CUDA(cudaMemsetAsync(keep.back()->p, 0, part.packed.size() * 4 + 256, c.stream));
      c.h2d(keep.back()->p, part.packed.data(), part.packed.size() * 4);
      seqs[r].packed = keep.back()->as<uint32_t>();
      keep.emplace_back(new DevMem(sizeof(int64_t) * (ns + 1)));
      c.h2d(keep.back()->p, part.starts.data(), sizeof(int64_t) * (ns + 1));
      seqs[r].starts = keep.back()->as<int64_t>();
      keep.emplace_back(new DevMem(sizeof(uint16_t) * (ns + 1)));
      c.h2d(keep.back()->p, part.mult.data(), sizeof(uint16_t) * ns);
      seqs[r].mult = keep.back()->as<uint16_t>();
      keep.emplace_back(new DevMem(sizeof(int64_t) * (ns + 1)));
      c.h2d(keep.back()->p, part.item_base.data(), sizeof(int64_t) * (ns + 1));
      seqs[r].item_base = keep.back()->as<int64_t>();
      seqs[r].nseq = ns;
      seqs[r].n_items = part.item_base.back();
    }
  }
  mg.seq2sdbg(d_edges, n_edges, seqs, k);
  write_sdbg_multi(mg, out_prefix);
  keep.clear();
}

}  // namespace mf
