// hostio.cu -- files in, files out, the device pipelines in between.  The on-disk layouts restate what
// megahit v1.2.9 reads and writes (EdgeWriter/EdgeIoMetadata, SdbgWriter/SdbgMeta, BinaryWriter, lib_info,
// ContigReader); the reference only fixes the file NAMES (assemble/assemble_wrapper.py:93-97,166-200,228-250).
#include "hostio.h"
#include <zlib.h>
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include "mfsdbg.h"
#include "sdbg_ser.cuh"

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
  if (fseeko(f, 0, SEEK_END) == 0) {   // a regular file: one allocation of its size (+1: the read that finds the end)
    const off_t sz = ftello(f);
    if (sz > 0) cap = (size_t)sz + 1;
  }
  rewind(f);
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
// MFSDBG_TRACE: host-clock split of a file-level call (input / device / output)
struct PhaseClock {
  const char *what;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  std::string line;
  explicit PhaseClock(const char *w) : what(w) {}
  void mark(const char *phase) {
    const auto t1 = std::chrono::steady_clock::now();
    char buf[64];
    snprintf(buf, sizeof buf, " %s %.1f ms", phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
    line += buf;
    t0 = t1;
  }
  ~PhaseClock() {
    if (getenv("MFSDBG_TRACE")) fprintf(stderr, "[mfsdbg] %s:%s\n", what, line.c_str());
  }
};

}  // namespace

// ---------------------------------------------------------------- buildlib
namespace {
// text bytes per file and chunk (pinned, two sets: one is read while one is packed); MFSDBG_TEXT_CHUNK overrides (tests)
static size_t text_chunk() {
  const char *v = getenv("MFSDBG_TEXT_CHUNK");
  return v && *v ? (size_t)std::max(1024ll, atoll(v)) : (size_t)64 << 20;   // 4 pinned buffers of this size: 192 MB each cost 0.2 s to pin
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
    // newlines of the chunk (a branch-free count the compiler vectorises: a memchr per 150-byte line cost as much as reading the
    // file); the cut goes behind the last one that completes a record, found by walking back over the lines of the partial record
    int64_t lines = 0;
    for (size_t i = 0; i < len; ++i) lines += buf[i] == '\n';
    size_t last_rec_end = 0;
    if (lines_per_rec && lines >= lines_per_rec) {
      size_t q = len;
      for (int64_t back = lines % lines_per_rec + 1; back > 0; --back) {
        const uint8_t *nl = (const uint8_t *)memrchr(buf, '\n', q);
        q = (size_t)(nl - buf);
      }
      last_rec_end = q + 1;
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
        if (nf == 2) {   // the mates' files side by side
          std::exception_ptr merr;
          std::thread mate([&] { try { st->len[1] = in[1]->next(st->buf[1].as<uint8_t>(), kTextChunk, &nr[1]); } catch (...) { merr = std::current_exception(); } });
          try {
            st->len[0] = in[0]->next(st->buf[0].as<uint8_t>(), kTextChunk, &nr[0]);
          } catch (...) {
            mate.join();
            throw;
          }
          mate.join();
          if (merr) std::rethrow_exception(merr);
        } else {
          st->len[0] = in[0]->next(st->buf[0].as<uint8_t>(), kTextChunk, &nr[0]);
        }
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

// ---------------------------------------------------------------- output streams
namespace {
// bytes per piece of an output stream (two pinned pieces: one crosses PCIe while the other is written); MFSDBG_IO_CHUNK overrides (tests)
static size_t io_chunk() {
  const char *v = getenv("MFSDBG_IO_CHUNK");
  const size_t b = v && *v ? (size_t)std::max(64ll, atoll(v)) : (size_t)32 << 20;
  return (b + 7) & ~(size_t)7;
}
// a byte stream cut into files at cumulative byte ends (files nobody's bytes reach stay empty)
struct SplitFiles {
  std::vector<File *> files;
  std::vector<int64_t> end;
  size_t cur = 0;
  int64_t pos = 0;
  void operator()(const uint8_t *p, size_t n) {
    while (n) {
      while (cur < end.size() && pos >= end[cur]) ++cur;
      if (cur >= end.size()) throw std::runtime_error("output stream is longer than its file table");
      const size_t take = (size_t)std::min<int64_t>((int64_t)n, end[cur] - pos);
      files[cur]->write(p, take);
      p += take;
      n -= take;
      pos += (int64_t)take;
    }
  }
};
// The device side fills one pinned piece (a synchronous copy on the context's stream) while a helper thread writes the other: the
// records never sit whole in pageable memory (round-2 files run: `.edges` 409 MB and `.sdbg` 206 MB went device -> zero-filled
// vector -> fwrite, one after the other).
class PipeWriter {
 public:
  PipeWriter(Ctx &c, size_t piece_bytes, std::function<void(const uint8_t *, size_t)> sink) : sink_(std::move(sink)) {
    for (int i = 0; i < 2; ++i) {
      c.io_pin[i].reserve(piece_bytes);
      buf_[i] = c.io_pin[i].as<uint8_t>();
    }
    th_ = std::thread([this] { run(); });
  }
  ~PipeWriter() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
  }
  // piece i, once whatever was submitted from it has been written
  uint8_t *piece(int i) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !busy_[i]; });
    if (err_) std::rethrow_exception(err_);
    return buf_[i];
  }
  void submit(int i, size_t n) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      busy_[i] = true;
      q_.push_back({i, n});
    }
    cv_.notify_all();
  }
  void finish() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return q_.empty() && !busy_[0] && !busy_[1]; });
      stop_ = true;
    }
    cv_.notify_all();
    th_.join();
    if (err_) std::rethrow_exception(err_);
  }

 private:
  void run() {
    for (;;) {
      std::pair<int, size_t> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        job = q_.front();
        q_.pop_front();
      }
      try {
        if (!err_) sink_(buf_[job.first], job.second);
      } catch (...) {
        std::lock_guard<std::mutex> lk(mu_);
        if (!err_) err_ = std::current_exception();
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        busy_[job.first] = false;
      }
      cv_.notify_all();
    }
  }
  std::function<void(const uint8_t *, size_t)> sink_;
  uint8_t *buf_[2] = {nullptr, nullptr};
  bool busy_[2] = {false, false};
  bool stop_ = false;
  std::deque<std::pair<int, size_t>> q_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::exception_ptr err_;
  std::thread th_;
};

// `bytes` of device memory through the two pinned pieces into `sink`
static void stream_device_bytes(Ctx &c, const void *dev, size_t bytes, std::function<void(const uint8_t *, size_t)> sink) {
  if (!bytes) return;
  const size_t chunk = io_chunk();
  PipeWriter pw(c, chunk, std::move(sink));
  int i = 0;
  for (size_t off = 0; off < bytes; off += chunk, i ^= 1) {
    const size_t n = std::min(chunk, bytes - off);
    uint8_t *h = pw.piece(i);
    c.d2h(h, (const uint8_t *)dev + off, n);
    pw.submit(i, n);
  }
  pw.finish();
}

// the graph's records as SdbgWriter's byte stream (sdbg_ser.cuh), piece by piece into `sink`; returns the stream's length
static int64_t stream_sdbg_bytes(Ctx &c, const SdbgView &g, std::function<void(const uint8_t *, size_t)> sink) {
  const int64_t n = g.n_items;
  if (n == 0) return 0;
  const int wt = g.words_tip;
  const int64_t ntiles = (n + kSerTile - 1) / kSerTile;
  DevMem d_lt(sizeof(uint32_t) * (size_t)ntiles), d_lb(sizeof(int64_t) * (size_t)(ntiles + 1)), d_tb(sizeof(int64_t) * (size_t)(ntiles + 1));
  k_ser_tile_counts<<<(unsigned)ntiles, kSerNT, 0, c.stream>>>(g.rec, n, d_lt.as<uint32_t>());
  MF_LAUNCH_CHECK();
  k_ser_scan<<<1, 1024, 0, c.stream>>>(d_lt.as<uint32_t>(), ntiles, d_lb.as<int64_t>(), d_tb.as<int64_t>());
  MF_LAUNCH_CHECK();
  c.launches += 2;
  std::vector<int64_t> lb((size_t)ntiles + 1), tb((size_t)ntiles + 1);
  c.d2h(lb.data(), d_lb.p, sizeof(int64_t) * lb.size());
  c.d2h(tb.data(), d_tb.p, sizeof(int64_t) * tb.size());
  if (tb[(size_t)ntiles] != g.n_tips) throw std::runtime_error("sdbg tip flags do not add up to the tip count");
  auto unit_at = [&](int64_t t) { return std::min<int64_t>(n, t * kSerTile) + lb[(size_t)t] + 2 * (int64_t)wt * tb[(size_t)t]; };
  const size_t chunk = io_chunk();
  const size_t tile_max = (size_t)kSerTile * (size_t)(2 + 2 * wt) * 2;   // a tile of large-multiplicity tips
  const size_t piece = chunk + tile_max;
  DevMem d_out(piece);
  PipeWriter pw(c, piece, std::move(sink));
  int i = 0;
  for (int64_t t0 = 0; t0 < ntiles; i ^= 1) {
    int64_t t1 = t0 + 1;
    while (t1 < ntiles && (size_t)(unit_at(t1 + 1) - unit_at(t0)) * 2 <= chunk) ++t1;
    const size_t bytes = (size_t)(unit_at(t1) - unit_at(t0)) * 2;
    uint8_t *h = pw.piece(i);
    k_ser_write<<<(unsigned)(t1 - t0), kSerNT, 0, c.stream>>>(g.rec, g.labels, n, wt, t0, d_lb.as<int64_t>(), d_tb.as<int64_t>(), unit_at(t0),
                                                             d_out.as<uint16_t>());
    MF_LAUNCH_CHECK();
    c.launches++;
    c.d2h(h, d_out.p, bytes);
    pw.submit(i, bytes);
    t0 = t1;
  }
  pw.finish();
  return unit_at(ntiles) * 2;
}
}  // namespace

// ---------------------------------------------------------------- edges files
static void write_edges(Ctx &c, const EdgesView &e, const std::string &prefix, int n_files) {
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
    info << b << ' ' << f << ' ' << foff[f] << ' ' << cnt << '\n';
    foff[f] += cnt;
    pos += cnt;
  }
  if (pos != e.n_edges) throw std::runtime_error("bucket counts do not add up to the edge count");
  // the sorted edges are one stream; file f holds the buckets bucket_file() gives it, i.e. a contiguous piece of it
  SplitFiles split;
  int64_t acc = 0;
  for (int f = 0; f < n_files; ++f) {
    acc += foff[f] * e.words * 4;
    split.files.push_back(files[f].get());
    split.end.push_back(acc);
  }
  stream_device_bytes(c, e.edges, (size_t)e.n_edges * e.words * 4, std::ref(split));
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
// where the edges of `<prefix>.edges.*` lie, in stream order: pieces of the files (sorted edges: bucket by bucket as the meta file
// lists them, neighbours merged; unsorted "iterative" edges: every file whole)
struct EdgeSeg {
  int file;
  int64_t off, len;   // bytes
};
struct EdgeFilesPlan {
  std::string prefix;
  int k = 0, words = 0, nfiles = 0;
  bool sorted = false;
  int64_t n = 0;
  std::vector<EdgeSeg> segs;
};
static int64_t file_size(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) throw IoError(errno_msg("cannot open", path));
  if (fseeko(f, 0, SEEK_END) != 0) { fclose(f); throw IoError(errno_msg("cannot seek in", path)); }
  const int64_t n = (int64_t)ftello(f);
  fclose(f);
  return n;
}
static EdgeFilesPlan plan_edges(const std::string &prefix) {
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
  EdgeFilesPlan e;
  e.prefix = prefix;
  e.k = (int)k; e.words = (int)words; e.sorted = sorted != 0; e.n = nedges; e.nfiles = (int)nfiles;
  std::vector<int64_t> fsize(nfiles);
  for (int f = 0; f < nfiles; ++f) fsize[f] = file_size(prefix + ".edges." + std::to_string(f));
  int64_t pos = 0;
  const int64_t rec = (int64_t)words * 4;
  if (e.sorted) {
    for (int b = 0; b < nbuckets; ++b) {
      long long bid, fid, off, cnt;
      if (!(is >> bid >> fid >> off >> cnt) || bid != b) throw IoError("Invalid format: bucket id not matched!");
      if (fid < 0 || cnt == 0) continue;
      if (off < 0 || cnt < 0) throw IoError("edges.info: negative offset or count");
      if (fid >= nfiles || (off + cnt) * rec > fsize[fid] || pos + cnt > nedges) throw IoError("edge file shorter than its meta says");
      if (!e.segs.empty() && e.segs.back().file == (int)fid && e.segs.back().off + e.segs.back().len == off * rec) e.segs.back().len += cnt * rec;
      else e.segs.push_back(EdgeSeg{(int)fid, off * rec, cnt * rec});
      pos += cnt;
    }
  } else {
    for (int f = 0; f < nfiles; ++f) {
      const int64_t cnt = fsize[f] / rec;
      if (pos + cnt > nedges) throw IoError("more edges on disk than the meta says");
      if (cnt) e.segs.push_back(EdgeSeg{f, 0, cnt * rec});
      pos += cnt;
    }
  }
  if (pos != nedges) throw IoError("edge count mismatch between meta and files");
  return e;
}
// the plan's bytes, in order, handed to `take(ptr, n)` in pieces of at most `piece` bytes read into buf[0], buf[1] alternately;
// `before(i)` runs before buffer i is overwritten
template <class Before, class Take>
static void read_edge_bytes(const EdgeFilesPlan &pl, uint8_t *const buf[2], size_t piece, Before &&before, Take &&take) {
  std::vector<std::unique_ptr<File>> files(pl.nfiles);
  int i = 0;
  for (const EdgeSeg &sg : pl.segs) {
    if (!files[sg.file]) files[sg.file].reset(new File(pl.prefix + ".edges." + std::to_string(sg.file), "rb"));
    File &f = *files[sg.file];
    if (fseeko(f.fp, (off_t)sg.off, SEEK_SET) != 0) throw IoError(errno_msg("cannot seek in", f.path));
    for (int64_t done = 0; done < sg.len;) {
      const size_t n = (size_t)std::min<int64_t>((int64_t)piece, sg.len - done);
      before(i);
      if (fread(buf[i], 1, n, f.fp) != n) throw IoError("read error on " + f.path);
      take(i, n);
      done += (int64_t)n;
      i ^= 1;
    }
  }
}
static HostEdges read_edges(const std::string &prefix) {
  const EdgeFilesPlan pl = plan_edges(prefix);
  HostEdges e;
  e.k = pl.k; e.words = pl.words; e.sorted = pl.sorted; e.n = pl.n;
  e.data.resize((size_t)pl.n * pl.words);
  uint8_t *dst = reinterpret_cast<uint8_t *>(e.data.data());
  size_t at = 0;
  std::vector<std::unique_ptr<File>> files(pl.nfiles);
  for (const EdgeSeg &sg : pl.segs) {
    if (!files[sg.file]) files[sg.file].reset(new File(pl.prefix + ".edges." + std::to_string(sg.file), "rb"));
    File &f = *files[sg.file];
    if (fseeko(f.fp, (off_t)sg.off, SEEK_SET) != 0) throw IoError(errno_msg("cannot seek in", f.path));
    if (fread(dst + at, 1, (size_t)sg.len, f.fp) != (size_t)sg.len) throw IoError("read error on " + f.path);
    at += (size_t)sg.len;
  }
  return e;
}
// the edges straight into device memory through the two pinned pieces (the copy of piece i runs while piece i+1 is read)
static void upload_edges(Ctx &c, const EdgeFilesPlan &pl, uint8_t *d_edges) {
  if (pl.n == 0) return;
  const size_t piece = io_chunk();
  for (HostBuf &b : c.io_pin) b.reserve(piece);
  uint8_t *const buf[2] = {c.io_pin[0].as<uint8_t>(), c.io_pin[1].as<uint8_t>()};
  cudaEvent_t ev[2];
  bool used[2] = {false, false};
  for (auto &e : ev) MF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  size_t at = 0;
  try {
    read_edge_bytes(
        pl, buf, piece,
        [&](int i) {
          if (used[i]) MF_CUDA(cudaEventSynchronize(ev[i]));
        },
        [&](int i, size_t n) {
          MF_CUDA(cudaMemcpyAsync(d_edges + at, buf[i], n, cudaMemcpyHostToDevice, c.stream));
          MF_CUDA(cudaEventRecord(ev[i], c.stream));
          used[i] = true;
          at += n;
        });
    MF_CUDA(cudaStreamSynchronize(c.stream));
  } catch (...) {
    cudaStreamSynchronize(c.stream);
    for (auto &e : ev) cudaEventDestroy(e);
    throw;
  }
  for (auto &e : ev) cudaEventDestroy(e);
}

void file_count(Ctx &c, const char *read_lib_file, int k, int min_count, const char *out_prefix, int n_files) {
  PhaseClock pc("count");
  ReadsView r;
  load_read_lib(c, read_lib_file, &r);
  pc.mark("read library");
  EdgesView e;
  std::vector<int64_t> counting(kNumBuckets, 0);
  c.begin_call();
  dev_count(c, r, k, min_count, &e, counting.data());
  c.end_call();
  pc.mark("device");
  write_edges(c, e, out_prefix, n_files);
  pc.mark("write edges");
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
  std::vector<std::unique_ptr<File>> files;
  for (int f = 0; f < n_files; ++f) files.emplace_back(new File(prefix + ".sdbg." + std::to_string(f), "wb"));
  std::ostringstream info;
  info << "k " << g.k << "\nwords_per_tip_label " << g.words_tip << "\nnum_buckets " << kNumBuckets << "\nnum_files " << n_files << '\n';
  std::vector<int64_t> foff(n_files, 0);
  // SdbgMeta: a bucket record nobody wrote to keeps bucket_id = kUninitializedBucketID = size_t(-1) and zeros elsewhere, and
  // the records are sorted by bucket_id before they are serialised -- the unused ones come LAST (recollection shared by the
  // oracle; the first run against a real megahit_core decides it)
  int64_t pos = 0, tpos = 0, large = 0, n_empty = 0;
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t items = c.sdbg_bucket_stats[(size_t)b * 3], tips = c.sdbg_bucket_stats[(size_t)b * 3 + 1],
                  lg = c.sdbg_bucket_stats[(size_t)b * 3 + 2];
    if (!items) { ++n_empty; continue; }
    const int f = bucket_file(b, n_files);
    info << b << ' ' << f << ' ' << foff[f] << ' ' << items << ' ' << tips << ' ' << lg << '\n';
    foff[f] += 2 * items + 2 * lg + 4 * (int64_t)g.words_tip * tips;   // SdbgWriter::Write: 16 bits, + 16 beyond 254, + the tip's label
    pos += items;
    tpos += tips;
    large += lg;
  }
  if (pos != g.n_items || tpos != g.n_tips) throw std::runtime_error("sdbg bucket statistics do not add up");
  for (int64_t i = 0; i < n_empty; ++i) info << "18446744073709551615 0 0 0 0 0\n";
  info << "item_count " << g.n_items << "\ntip_count " << g.n_tips << "\nlarge_mul_count " << large << '\n';
  // the records are serialised on the device (sdbg_ser.cuh); file f holds a contiguous piece of that stream
  SplitFiles split;
  int64_t acc = 0;
  for (int f = 0; f < n_files; ++f) {
    acc += foff[f];
    split.files.push_back(files[f].get());
    split.end.push_back(acc);
  }
  if (stream_sdbg_bytes(c, g, std::ref(split)) != acc) throw std::runtime_error("sdbg stream length does not match the bucket statistics");
  for (auto &f : files) f->close();
  File fi(prefix + ".sdbg_info", "w");
  const std::string s = info.str();
  fi.write(s.data(), s.size());
  fi.close();
}

void file_seq2sdbg(Ctx &c, int k, int k_from, const char *input_prefix, const char *contig, const char *bubble,
                   const char *addi_contig, const char *local_contig, const char *out_prefix, int n_files) {
  PhaseClock pc("seq2sdbg");
  EdgeFilesPlan he;
  if (input_prefix && *input_prefix) {
    he = plan_edges(input_prefix);
    if (he.k != k) throw IoError("edges were built for k=" + std::to_string(he.k) + ", not " + std::to_string(k));
  }
  HostSeqs hs;
  if (contig && *contig) add_contigs(&hs, contig, k, true, k_from, k);
  if (bubble && *bubble) add_contigs(&hs, bubble, k, true, k_from, k);
  if (addi_contig && *addi_contig) add_contigs(&hs, addi_contig, k, false, 0, 0);
  if (local_contig && *local_contig) add_contigs(&hs, local_contig, k, false, 0, 0);
  const int nseq = (int)hs.mult.size();
  pc.mark("contigs");
  DevMem d_edges((size_t)he.n * he.words * 4 + 64), d_packed(hs.packed.size() * 4 + 256), d_starts(sizeof(int64_t) * (nseq + 1)),
      d_mult(sizeof(uint16_t) * (nseq + 1)), d_ibase(sizeof(int64_t) * (nseq + 1));
  upload_edges(c, he, d_edges.as<uint8_t>());
  pc.mark("read + upload edges");
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
  pc.mark("upload contigs");
  SdbgView g;
  c.begin_call();
  dev_seq2sdbg(c, d_edges.as<uint32_t>(), he.n, sv, k, 0, &g);
  c.end_call();
  pc.mark("device");
  write_sdbg(c, g, out_prefix, n_files);
  pc.mark("write sdbg");
}

void file_read2sdbg(Ctx &c, const char *read_lib_file, int k, int min_count, const char *out_prefix, int n_files) {
  PhaseClock pc("read2sdbg");
  ReadsView r;
  load_read_lib(c, read_lib_file, &r);
  pc.mark("read library");
  EdgesView e;
  SdbgView g;
  c.begin_call();
  dev_count(c, r, k, min_count, &e, nullptr);
  dev_seq2sdbg(c, e.edges, e.n_edges, SeqsView{}, k, 1, &g);
  c.end_call();
  pc.mark("device");
  write_sdbg(c, g, out_prefix, n_files);
  pc.mark("write sdbg");
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
    File f(prefix + ".edges." + std::to_string(r), "wb");
    SplitFiles split;
    split.files.push_back(&f);
    split.end.push_back(e.n_edges * e.words * 4);
    stream_device_bytes(c, e.edges, (size_t)e.n_edges * e.words * 4, std::ref(split));
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
  for (int r = 0; r < G; ++r) {
    Ctx &c = mg.ctx(r);
    const SdbgView &g = mg.sdbg(r);
    MF_CUDA(cudaSetDevice(c.device));
    File f(prefix + ".sdbg." + std::to_string(r), "wb");
    int64_t pos = 0, tpos = 0, foff = 0;
    for (int b = 0; b < kNumBuckets; ++b) {
      const int64_t items = c.sdbg_bucket_stats[(size_t)b * 3], tips = c.sdbg_bucket_stats[(size_t)b * 3 + 1],
                    lg = c.sdbg_bucket_stats[(size_t)b * 3 + 2];
      if (!items) continue;
      if (!rows[b].empty()) throw std::runtime_error("sdbg bucket " + std::to_string(b) + " is held by two GPUs");
      rows[b] = std::to_string(b) + ' ' + std::to_string(r) + ' ' + std::to_string(foff) + ' ' + std::to_string(items) + ' ' +
                std::to_string(tips) + ' ' + std::to_string(lg) + '\n';
      foff += 2 * items + 2 * lg + 4 * (int64_t)g.words_tip * tips;
      pos += items;
      tpos += tips;
      n_large += lg;
    }
    if (pos != g.n_items || tpos != g.n_tips) throw std::runtime_error("sdbg bucket statistics do not add up");
    SplitFiles split;
    split.files.push_back(&f);
    split.end.push_back(foff);
    if (stream_sdbg_bytes(c, g, std::ref(split)) != foff) throw std::runtime_error("sdbg stream length does not match the bucket statistics");
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
