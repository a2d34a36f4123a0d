// pack.cu -- K1: FASTQ/FASTA text in HBM -> 2-bit packed reads (what `megahit_core buildlib` produces), plus the
// conversions between the contiguous device layout and the on-disk reads.lib.bin record stream.
//   k_newline_count / k_newline_fill : line index (one thread per 16 text bytes, block scan)
//   k_segments_*                     : per record, the kept segment(s) under the N policy
//   k_pack_words                     : one thread per output word (16 bases)
// N policy (FastxReader::TrimN of megahit v1.2.9): a read keeps its first N-free segment; MFSDBG_N_SPLIT instead makes
// every N-free segment a read of its own.  Base codes: "ACGTNacgtn" -> 0123201232, anything else 0 (SequencePackage).
#include "engine.cuh"
#include "hostio.h"
#include "mfsdbg.h"

namespace mf {

void device_excl_scan_i64(Ctx &c, const int64_t *v, int64_t n, int64_t *out);

__device__ __forceinline__ uint32_t dna_code(uint8_t ch) {
  switch (ch) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': case 'N': case 'n': return 2;
    case 'T': case 't': return 3;
    default: return 0;
  }
}
__device__ __forceinline__ bool is_n(uint8_t ch) { return ch == 'N' || ch == 'n'; }

constexpr int kNlBytesPerThread = 16;
constexpr int kNlThreads = 256;
constexpr int kNlBytesPerBlock = kNlBytesPerThread * kNlThreads;

__global__ void k_newline_count(const uint8_t *__restrict__ text, int64_t n, int64_t *__restrict__ block_cnt) {
  __shared__ int s[kNlThreads / 32];
  const int64_t b = (int64_t)blockIdx.x * kNlBytesPerBlock + (int64_t)threadIdx.x * kNlBytesPerThread;
  int cnt = 0;
  for (int i = 0; i < kNlBytesPerThread; ++i) cnt += (b + i < n) && text[b + i] == '\n';
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < kNlThreads / 32; ++i) t += s[i];
    block_cnt[blockIdx.x] = t;
  }
}
// line_start[0] = 0 ; line_start[j+1] = position after the j-th newline
__global__ void k_newline_fill(const uint8_t *__restrict__ text, int64_t n, const int64_t *__restrict__ block_off,
                               int64_t *__restrict__ line_start) {
  __shared__ int s[kNlThreads / 32 + 1];
  const int64_t b = (int64_t)blockIdx.x * kNlBytesPerBlock + (int64_t)threadIdx.x * kNlBytesPerThread;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cnt = 0;
  for (int i = 0; i < kNlBytesPerThread; ++i) cnt += (b + i < n) && text[b + i] == '\n';
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s[warp] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < kNlThreads / 32; ++i) { int v = s[i]; s[i] = run; run += v; }
  }
  __syncthreads();
  int64_t rank = block_off[blockIdx.x] + s[warp] + inc - cnt;
  for (int i = 0; i < kNlBytesPerThread; ++i)
    if (b + i < n && text[b + i] == '\n') line_start[++rank] = b + i + 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
}

struct TextSet {
  const uint8_t *t[2];
  int64_t n[2];
  const int64_t *line_start[2];   // device
  int64_t n_lines[2];
  int lines_per_rec;              // 4 = FASTQ, 2 = FASTA
  int n_texts;
};
constexpr int64_t kTextIdBit = (int64_t)1 << 62;

__device__ __forceinline__ void seq_line(const TextSet &ts, int f, int64_t rec, const uint8_t **p, int64_t *len) {
  const int64_t li = rec * ts.lines_per_rec + 1;
  int64_t b = ts.line_start[f][li];
  int64_t e = ts.line_start[f][li + 1] - 1;   // line_start[n_lines] exists (real or sentinel)
  if (e > ts.n[f]) e = ts.n[f];
  if (e > b && ts.t[f][e - 1] == '\r') --e;
  *p = ts.t[f] + b;
  *len = e - b;
}
// record r of the interleaved order -> (file, record)
__device__ __forceinline__ void which(const TextSet &ts, int64_t r, int *f, int64_t *rec) {
  if (ts.n_texts == 2) { *f = (int)(r & 1); *rec = r >> 1; } else { *f = 0; *rec = r; }
}
// validate headers + count segments per record
__global__ void k_segments_count(TextSet ts, int64_t n_rec_total, int n_policy, int64_t *__restrict__ seg_cnt, int *err) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec_total) return;
  int f; int64_t rec;
  which(ts, r, &f, &rec);
  const uint8_t h = ts.t[f][ts.line_start[f][rec * ts.lines_per_rec]];
  if (ts.lines_per_rec == 4) {
    const uint8_t p = ts.t[f][ts.line_start[f][rec * 4 + 2]];
    if (h != '@' || p != '+') atomicExch(err, 1);
  } else if (h != '>') {
    atomicExch(err, 1);
  }
  if (n_policy != MFSDBG_N_SPLIT) { seg_cnt[r] = 1; return; }
  const uint8_t *s; int64_t len;
  seq_line(ts, f, rec, &s, &len);
  int64_t c = 0; bool in = false;
  for (int64_t i = 0; i < len; ++i) {
    bool nn = is_n(s[i]);
    if (!nn && !in) ++c;
    in = !nn;
  }
  seg_cnt[r] = c ? c : 1;
}
__global__ void k_segments_fill(TextSet ts, int64_t n_rec_total, int n_policy, const int64_t *__restrict__ seg_base,
                                int64_t *__restrict__ seg_off, int64_t *__restrict__ seg_len, int *max_len) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec_total) return;
  int f; int64_t rec;
  which(ts, r, &f, &rec);
  const uint8_t *s; int64_t len;
  seq_line(ts, f, rec, &s, &len);
  const int64_t text_off = (s - ts.t[f]) | (f ? kTextIdBit : 0);
  int64_t o = seg_base[r];
  int mx = 0;
  if (n_policy != MFSDBG_N_SPLIT) {
    int64_t b = len, i;
    for (i = 0; i < len; ++i) {
      if (is_n(s[i])) { if (b < len) break; }
      else if (b == len) b = i;
    }
    int64_t e = i;
    if (b > e) b = e;
    seg_off[o] = text_off + b;
    seg_len[o] = e - b;
    mx = (int)(e - b);
  } else {
    int64_t i = 0; bool any = false;
    while (i < len) {
      while (i < len && is_n(s[i])) ++i;
      int64_t b = i;
      while (i < len && !is_n(s[i])) ++i;
      if (i > b) { seg_off[o] = text_off + b; seg_len[o] = i - b; if (i - b > mx) mx = (int)(i - b); ++o; any = true; }
    }
    if (!any) { seg_off[o] = text_off; seg_len[o] = 0; }
  }
  if (mx > 0) atomicMax(max_len, mx);
}
__global__ void k_pack_words(TextSet ts, int64_t n_reads, const int64_t *__restrict__ starts, const int64_t *__restrict__ seg_off,
                             int64_t n_bases, uint32_t *__restrict__ words, int64_t n_words) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  int64_t g = w * 16;
  if (g >= n_bases) { words[w] = 0; return; }
  int64_t lo = 0, hi = n_reads;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (starts[mid] <= g) lo = mid; else hi = mid;
  }
  int64_t r = lo, rs = starts[r], re = starts[r + 1];
  const uint8_t *src = ts.t[(seg_off[r] & kTextIdBit) ? 1 : 0] + (seg_off[r] & ~kTextIdBit);
  uint32_t out = 0;
  for (int i = 0; i < 16 && g < n_bases; ++i, ++g) {
    while (g >= re) {
      ++r;
      rs = starts[r];
      re = starts[r + 1];
      src = ts.t[(seg_off[r] & kTextIdBit) ? 1 : 0] + (seg_off[r] & ~kTextIdBit);
    }
    out |= dna_code(src[g - rs]) << (30 - 2 * i);
  }
  words[w] = out;
}

static void index_lines(Ctx &c, const uint8_t *text, int64_t n, DevBuf *line_buf, int64_t *n_lines) {
  const int64_t nb = std::max<int64_t>(1, div_ceil64(n, kNlBytesPerBlock));
  int64_t *d_cnt = nullptr;
  MF_CUDA(cudaMalloc(&d_cnt, sizeof(int64_t) * (2 * nb + 2)));
  k_newline_count<<<(unsigned)nb, kNlThreads, 0, c.stream>>>(text, n, d_cnt);
  MF_LAUNCH_CHECK();
  c.launches++;
  device_excl_scan_i64(c, d_cnt, nb, d_cnt + nb);
  int64_t total = 0;
  c.d2h(&total, d_cnt + 2 * nb, 8);
  uint8_t last = '\n';
  if (n > 0) c.d2h(&last, text + n - 1, 1);
  *n_lines = total + (n > 0 && last != '\n');
  line_buf->reserve(sizeof(int64_t) * (total + 3));
  k_newline_fill<<<(unsigned)nb, kNlThreads, 0, c.stream>>>(text, n, d_cnt + nb, line_buf->as<int64_t>());
  MF_LAUNCH_CHECK();
  c.launches++;
  // sentinel entries so line_start[n_lines] and [n_lines+1] are readable
  int64_t sent[2] = {n + 1, n + 1};
  if (last != '\n' || n == 0) MF_CUDA(cudaMemcpyAsync(line_buf->as<int64_t>() + total + 1, sent, 16, cudaMemcpyHostToDevice, c.stream));
  else MF_CUDA(cudaMemcpyAsync(line_buf->as<int64_t>() + total + 1, sent, 8, cudaMemcpyHostToDevice, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  cudaFree(d_cnt);
}

void dev_pack_fastq(Ctx &c, const uint8_t *const *texts, const int64_t *n_bytes, int n_texts, int n_policy, ReadsView *out,
                    int *max_len) {
  if (n_texts < 1 || n_texts > 2) throw std::invalid_argument("pack: 1 or 2 texts");
  TextSet ts{};
  ts.n_texts = n_texts;
  DevBuf lines[2];
  uint8_t first = 0;
  if (n_bytes[0] > 0) c.d2h(&first, texts[0], 1);
  ts.lines_per_rec = first == '>' ? 2 : 4;
  int64_t n_rec[2] = {0, 0};
  for (int f = 0; f < n_texts; ++f) {
    ts.t[f] = texts[f];
    ts.n[f] = n_bytes[f];
    index_lines(c, texts[f], n_bytes[f], &lines[f], &ts.n_lines[f]);
    ts.line_start[f] = lines[f].as<int64_t>();
    if (ts.n_lines[f] % ts.lines_per_rec) {
      for (auto &l : lines) l.release();
      throw IoError("read file is not made of whole single-line FASTQ/FASTA records (multi-line records are not supported)");
    }
    n_rec[f] = ts.n_lines[f] / ts.lines_per_rec;
  }
  if (n_texts == 2 && n_rec[0] != n_rec[1]) {
    for (auto &l : lines) l.release();
    throw IoError("paired files have different numbers of reads");
  }
  const int64_t n_rec_total = n_rec[0] + n_rec[1];
  int64_t *d_cnt = nullptr, *d_base = nullptr;
  int *d_flags = nullptr;
  MF_CUDA(cudaMalloc(&d_cnt, sizeof(int64_t) * std::max<int64_t>(n_rec_total, 1)));
  MF_CUDA(cudaMalloc(&d_base, sizeof(int64_t) * (n_rec_total + 1)));
  MF_CUDA(cudaMalloc(&d_flags, sizeof(int) * 2));
  MF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * 2, c.stream));
  if (n_rec_total > 0) {
    k_segments_count<<<(unsigned)div_ceil64(n_rec_total, 128), 128, 0, c.stream>>>(ts, n_rec_total, n_policy, d_cnt, d_flags);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  device_excl_scan_i64(c, d_cnt, n_rec_total, d_base);
  int64_t n_reads = 0;
  int flags[2];
  c.d2h(&n_reads, d_base + n_rec_total, 8);
  c.d2h(flags, d_flags, sizeof flags);
  if (flags[0]) {
    cudaFree(d_cnt); cudaFree(d_base); cudaFree(d_flags);
    for (auto &l : lines) l.release();
    throw IoError("malformed FASTQ/FASTA record (header or '+' line not where a single-line record puts it)");
  }
  int64_t *d_off = nullptr, *d_len = nullptr;
  MF_CUDA(cudaMalloc(&d_off, sizeof(int64_t) * std::max<int64_t>(n_reads, 1)));
  MF_CUDA(cudaMalloc(&d_len, sizeof(int64_t) * std::max<int64_t>(n_reads, 1)));
  if (n_rec_total > 0) {
    k_segments_fill<<<(unsigned)div_ceil64(n_rec_total, 128), 128, 0, c.stream>>>(ts, n_rec_total, n_policy, d_base, d_off, d_len,
                                                                                d_flags + 1);
    MF_LAUNCH_CHECK();
    c.launches++;
  }
  c.pack_starts.reserve(sizeof(int64_t) * (n_reads + 2));
  int64_t *d_starts = c.pack_starts.as<int64_t>();
  device_excl_scan_i64(c, d_len, n_reads, d_starts);
  int64_t n_bases = 0;
  c.d2h(&n_bases, d_starts + n_reads, 8);
  c.d2h(flags, d_flags, sizeof flags);
  const int64_t n_words = ((n_bases + 15) >> 4) + 16;
  c.pack_words.reserve((size_t)n_words * 4);
  k_pack_words<<<(unsigned)div_ceil64(n_words, 256), 256, 0, c.stream>>>(ts, n_reads, d_starts, d_off, n_bases,
                                                                        c.pack_words.as<uint32_t>(), n_words);
  MF_LAUNCH_CHECK();
  c.launches++;
  MF_CUDA(cudaStreamSynchronize(c.stream));
  cudaFree(d_cnt); cudaFree(d_base); cudaFree(d_flags); cudaFree(d_off); cudaFree(d_len);
  for (auto &l : lines) l.release();
  out->packed = c.pack_words.as<uint32_t>();
  out->starts = d_starts;
  out->n_reads = n_reads;
  out->n_bases = n_bases;
  if (max_len) *max_len = flags[1];
}

// ---- reads.lib.bin record stream <-> contiguous device layout --------------------------------
// BinaryWriter: per read uint32 length then ceil(len/16) packed words.
__global__ void k_bin_sizes(const int64_t *starts, int64_t n_reads, int64_t *rec_words) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_reads) rec_words[r] = 1 + ((starts[r + 1] - starts[r] + 15) >> 4);
}
__global__ void k_bin_stream(const uint32_t *__restrict__ packed, const int64_t *__restrict__ starts, int64_t n_reads,
                             const int64_t *__restrict__ rec_off, uint32_t *__restrict__ stream) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const int64_t s = starts[r], len = starts[r + 1] - s;
  uint32_t *dst = stream + rec_off[r];
  dst[0] = (uint32_t)len;
  const int nw = (int)((len + 15) >> 4);
  for (int j = 0; j < nw; ++j) {
    const int64_t g = s + 16 * (int64_t)j;
    const int64_t wi = g >> 4;
    const int sh = (int)(g & 15) * 2;
    uint32_t v = __funnelshift_l(packed[wi + 1], packed[wi], sh);
    const int64_t left = len - 16 * (int64_t)j;
    if (left < 16) v &= 0xffffffffu << (32 - 2 * (int)left);
    dst[1 + j] = v;
  }
}
__global__ void k_unbin(const uint32_t *__restrict__ stream, const int64_t *__restrict__ rec_off, const int64_t *__restrict__ starts,
                        int64_t n_reads, uint32_t *__restrict__ packed) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const int64_t s = starts[r], len = starts[r + 1] - s;
  const uint32_t *src = stream + rec_off[r] + 1;
  const int nw = (int)((len + 15) >> 4);
  for (int j = 0; j < nw; ++j) {
    uint32_t v = src[j];
    const int64_t left = len - 16 * (int64_t)j;
    if (left < 16) v &= 0xffffffffu << (32 - 2 * (int)left);
    const int64_t g = s + 16 * (int64_t)j;
    const int64_t wi = g >> 4;
    const int sh = (int)(g & 15) * 2;
    atomicOr(packed + wi, v >> sh);
    if (sh) atomicOr(packed + wi + 1, v << (32 - sh));
  }
}

// device reads -> host .bin stream (returned as words)
void reads_to_bin_stream(Ctx &c, const ReadsView &r, std::vector<uint32_t> *stream) {
  stream->clear();
  if (r.n_reads == 0) return;
  int64_t *d_sz = nullptr, *d_off = nullptr;
  MF_CUDA(cudaMalloc(&d_sz, sizeof(int64_t) * r.n_reads));
  MF_CUDA(cudaMalloc(&d_off, sizeof(int64_t) * (r.n_reads + 1)));
  k_bin_sizes<<<(unsigned)div_ceil64(r.n_reads, 256), 256, 0, c.stream>>>(r.starts, r.n_reads, d_sz);
  MF_LAUNCH_CHECK();
  c.launches++;
  device_excl_scan_i64(c, d_sz, r.n_reads, d_off);
  int64_t total = 0;
  c.d2h(&total, d_off + r.n_reads, 8);
  uint32_t *d_stream = nullptr;
  MF_CUDA(cudaMalloc(&d_stream, (size_t)total * 4));
  k_bin_stream<<<(unsigned)div_ceil64(r.n_reads, 256), 256, 0, c.stream>>>(r.packed, r.starts, r.n_reads, d_off, d_stream);
  MF_LAUNCH_CHECK();
  c.launches++;
  stream->resize((size_t)total);
  MF_CUDA(cudaMemcpyAsync(stream->data(), d_stream, (size_t)total * 4, cudaMemcpyDeviceToHost, c.stream));
  MF_CUDA(cudaStreamSynchronize(c.stream));
  cudaFree(d_sz); cudaFree(d_off); cudaFree(d_stream);
}
// host .bin stream -> device reads (kept in ctx.pack_words / ctx.pack_starts)
void bin_stream_to_reads(Ctx &c, const uint32_t *stream, int64_t n_words_stream, ReadsView *out) {
  std::vector<int64_t> rec_off, starts;
  starts.push_back(0);
  int64_t p = 0;
  while (p < n_words_stream) {
    const uint32_t len = stream[p];
    const int64_t nw = (len + 15) >> 4;
    if (p + 1 + nw > n_words_stream) throw IoError("truncated read library (.bin)");
    rec_off.push_back(p);
    starts.push_back(starts.back() + len);
    p += 1 + nw;
  }
  const int64_t n_reads = (int64_t)rec_off.size(), n_bases = starts.back();
  const int64_t n_words = ((n_bases + 15) >> 4) + 16;
  c.pack_words.reserve((size_t)n_words * 4);
  c.pack_starts.reserve(sizeof(int64_t) * (n_reads + 2));
  MF_CUDA(cudaMemsetAsync(c.pack_words.p, 0, (size_t)n_words * 4, c.stream));
  MF_CUDA(cudaMemcpyAsync(c.pack_starts.p, starts.data(), sizeof(int64_t) * (n_reads + 1), cudaMemcpyHostToDevice, c.stream));
  if (n_reads > 0) {
    uint32_t *d_stream = nullptr;
    int64_t *d_off = nullptr;
    MF_CUDA(cudaMalloc(&d_stream, (size_t)n_words_stream * 4));
    MF_CUDA(cudaMalloc(&d_off, sizeof(int64_t) * n_reads));
    MF_CUDA(cudaMemcpyAsync(d_stream, stream, (size_t)n_words_stream * 4, cudaMemcpyHostToDevice, c.stream));
    MF_CUDA(cudaMemcpyAsync(d_off, rec_off.data(), sizeof(int64_t) * n_reads, cudaMemcpyHostToDevice, c.stream));
    k_unbin<<<(unsigned)div_ceil64(n_reads, 256), 256, 0, c.stream>>>(d_stream, d_off, c.pack_starts.as<int64_t>(), n_reads,
                                                                     c.pack_words.as<uint32_t>());
    MF_LAUNCH_CHECK();
    c.launches++;
    MF_CUDA(cudaStreamSynchronize(c.stream));
    cudaFree(d_stream); cudaFree(d_off);
  }
  MF_CUDA(cudaStreamSynchronize(c.stream));
  out->packed = c.pack_words.as<uint32_t>();
  out->starts = c.pack_starts.as<int64_t>();
  out->n_reads = n_reads;
  out->n_bases = n_bases;
}

}  // namespace mf
