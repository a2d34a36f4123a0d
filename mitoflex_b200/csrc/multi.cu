// multi.cu -- several GPUs of one box behind the calls the reference actually makes (ONE process per sub-command,
// /root/reference/assemble/assemble_wrapper.py:224,258): one host thread and one context per device, buffers mapped by peer
// access (no IPC, no torch.distributed).  The data path is the one mitoflex_b200/dist.py drives from Python, stage by stage:
//
//   count : reads split by index -> prefix histogram per GPU -> owners = contiguous prefix-bin ranges balanced on the global
//           histogram -> k_reads_scatter stores every key straight into its owner's HBM over NVLink (fused partition + exchange)
//           -> every GPU finishes a disjoint, contiguous key range
//   sdbg  : items of the local edges (+ this GPU's share of the contigs) -> the same exchange on item prefixes -> finish
//
// Rank order = key order, so rank r's output is a contiguous piece of the global streams: it is written as <prefix>.edges.<r> /
// <prefix>.sdbg.<r> and the meta file maps every 16-bit bucket to (file r, offset, count) -- the reference's own multi-file
// layout (one file per worker thread there, one per GPU here).
#include <condition_variable>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>
#include "hostio.h"

namespace mf {

namespace {

constexpr int kMultiL1Bits = 6;   // 64 exchange bins, as in dist.py l1_bits_for() (256 bins at 8 GPUs measured 2.8x slower)

struct RankAborted : std::runtime_error {
  RankAborted() : std::runtime_error("another GPU's thread failed") {}
};
// A rank that fails aborts the barrier: every waiter (now or later) leaves with RankAborted instead of deadlocking.
class Barrier {
 public:
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    if (aborted_) throw RankAborted();
    const int gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen != gen_ || aborted_; });
      if (gen == gen_) throw RankAborted();
    }
  }
  void abort() {
    std::lock_guard<std::mutex> lk(m_);
    aborted_ = true;
    cv_.notify_all();
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
  bool aborted_ = false;
};

// contiguous bin ranges [bounds[r], bounds[r+1]) per rank, balanced on the global histogram (dist.py assign_owners)
std::vector<int64_t> assign_owners(const std::vector<std::vector<int64_t>> &H, int world) {
  const int nb = (int)H[0].size();
  std::vector<int64_t> csum(nb + 1, 0);
  for (int b = 0; b < nb; ++b) {
    int64_t t = 0;
    for (int s = 0; s < world; ++s) t += H[s][b];
    csum[b + 1] = csum[b] + t;
  }
  const int64_t total = csum[nb];
  std::vector<int64_t> bounds(world + 1, 0);
  bounds[world] = nb;
  for (int r = 1; r < world; ++r) {
    const int64_t target = total * r / world;
    int64_t b = std::lower_bound(csum.begin(), csum.end(), target) - csum.begin();   // first boundary reaching the target
    bounds[r] = std::min<int64_t>(std::max<int64_t>(b, bounds[r - 1]), nb);
  }
  return bounds;
}
struct ExchangePlan {
  std::vector<int64_t> bounds;
  int lo = 0, hi = 0, n_segs = 1;
  int64_t n_recv = 0, n_send_remote = 0;
  std::vector<int64_t> chunk_start, chunk_size;   // receive buffer: source-major, bins ascending (dist.py exchange_plan)
  std::vector<int32_t> chunk_seg;
};
ExchangePlan exchange_plan(const std::vector<std::vector<int64_t>> &H, int rank) {
  const int world = (int)H.size();
  ExchangePlan p;
  p.bounds = assign_owners(H, world);
  p.lo = (int)p.bounds[rank];
  p.hi = (int)p.bounds[rank + 1];
  p.n_segs = std::max(p.hi - p.lo, 1);
  int64_t off = 0;
  for (int s = 0; s < world; ++s)
    for (int b = p.lo; b < p.hi; ++b) {
      const int64_t sz = H[s][b];
      if (sz > 0) {
        p.chunk_start.push_back(off);
        p.chunk_size.push_back(sz);
        p.chunk_seg.push_back(b - p.lo);
      }
      off += sz;
    }
  p.n_recv = off;
  for (int b = 0; b < (int)H[rank].size(); ++b)
    if (b < p.lo || b >= p.hi) p.n_send_remote += H[rank][b];
  return p;
}
// byte address at which this rank's records of every prefix bin land in their owner's receive buffer (dist.py peer_bin_bases)
std::vector<unsigned long long> peer_bin_bases(const std::vector<std::vector<int64_t>> &H, const std::vector<int64_t> &bounds, int rank,
                                               const std::vector<void *> &peer_ptrs, int rec_bytes) {
  const int world = (int)H.size(), nb = (int)H[0].size();
  std::vector<unsigned long long> out(nb, 0);
  for (int r = 0; r < world; ++r) {
    const int lo = (int)bounds[r], hi = (int)bounds[r + 1];
    int64_t base = 0;   // lower-ranked sources come first in r's buffer
    for (int s = 0; s < rank; ++s)
      for (int b = lo; b < hi; ++b) base += H[s][b];
    int64_t within = 0;
    for (int b = lo; b < hi; ++b) {
      out[b] = (unsigned long long)(uintptr_t)peer_ptrs[r] + (unsigned long long)(base + within) * (unsigned long long)rec_bytes;
      within += H[rank][b];
    }
  }
  return out;
}

struct Shared {   // what the rank threads share
  int world = 1;
  std::unique_ptr<Barrier> bar;
  std::vector<std::vector<int64_t>> H;     // [world][bins] histograms of the current exchange
  std::vector<void *> ptrs;                // receive buffers of the current exchange
  std::vector<std::exception_ptr> err;
};

}  // namespace

struct MultiRank {
  int rank = 0, device = 0;
  std::unique_ptr<Ctx> ctx;
  DevBuf key_buf, item_buf, scratch, items, small;
  EdgesView edges;
  SdbgView sdbg;
  std::vector<int64_t> counting;
  ~MultiRank() {
    if (ctx) cudaSetDevice(device);
    for (DevBuf *b : {&key_buf, &item_buf, &scratch, &items, &small}) b->release();
  }
};

// One exchange: histogram `hist_dev` (already computed on this rank) -> plan -> receive buffer -> scatter by `scatter` ->
// barrier.  Returns the plan; the records are in `recv` afterwards.
template <class Scatter>
static ExchangePlan exchange(Shared &sh, MultiRank &me, const unsigned long long *hist_dev, int rec_words, DevBuf &recv, Scatter &&scatter) {
  Ctx &c = *me.ctx;
  const int nb = 1 << kMultiL1Bits;
  std::vector<unsigned long long> h(nb);
  c.d2h(h.data(), hist_dev, sizeof(unsigned long long) * nb);
  sh.H[me.rank].assign(h.begin(), h.end());
  sh.bar->wait();
  ExchangePlan plan = exchange_plan(sh.H, me.rank);
  recv.reserve((size_t)std::max<int64_t>(plan.n_recv, 1) * rec_words * 4 + 256);
  sh.ptrs[me.rank] = recv.p;
  sh.bar->wait();
  const std::vector<unsigned long long> bases = peer_bin_bases(sh.H, plan.bounds, me.rank, sh.ptrs, rec_words * 4);
  unsigned long long *d_bases = me.small.as<unsigned long long>() + nb;
  c.h2d(d_bases, bases.data(), sizeof(unsigned long long) * nb);
  scatter(d_bases);
  MF_CUDA(cudaStreamSynchronize(c.stream));
  sh.bar->wait();   // every rank's stores have landed
  return plan;
}

static void rank_count(Shared &sh, MultiRank &me, const ReadsView &reads, int k, int min_count, bool want_counting) {
  Ctx &c = *me.ctx;
  const int nb = 1 << kMultiL1Bits, Wk = words_key(k);
  me.small.reserve(sizeof(unsigned long long) * 2 * nb + 256);
  unsigned long long *d_hist = me.small.as<unsigned long long>();
  dev_count_hist(c, reads, k, kMultiL1Bits, d_hist);
  ExchangePlan plan = exchange(sh, me, d_hist, Wk, me.key_buf, [&](const unsigned long long *d_bases) {
    dev_count_scatter(c, reads, k, kMultiL1Bits, nullptr, nullptr, -1, d_bases);
  });
  me.scratch.reserve((size_t)(std::max<int64_t>(plan.n_recv, 1) + 16) * Wk * 4);
  if (want_counting) me.counting.assign(kNumBuckets, 0);
  dev_count_finish(c, me.key_buf.as<uint32_t>(), me.scratch.as<uint32_t>(), plan.n_recv, plan.chunk_start.data(), plan.chunk_size.data(),
                   plan.chunk_seg.data(), (int)plan.chunk_start.size(), plan.n_segs, k, kMultiL1Bits, min_count, &me.edges,
                   want_counting ? me.counting.data() : nullptr);
  me.scratch.release();
  me.key_buf.release();
}
static void rank_sdbg(Shared &sh, MultiRank &me, const uint32_t *edges, int64_t n_edges, const SeqsView &seqs, int k, int tip_mode) {
  Ctx &c = *me.ctx;
  const int nb = 1 << kMultiL1Bits, Wi = words_item(k);
  me.small.reserve(sizeof(unsigned long long) * 2 * nb + 256);
  unsigned long long *d_hist = me.small.as<unsigned long long>();
  const int64_t n_items = 6 * n_edges + seqs.n_items;
  me.items.reserve((size_t)std::max<int64_t>(n_items, 1) * Wi * 4 + 256);
  const int64_t got = dev_sdbg_items_seqs(c, edges, n_edges, seqs, k, me.items.as<uint32_t>());
  if (got != n_items) throw std::runtime_error("multi-GPU sdbg: generated items disagree with their count");
  dev_records_hist(c, me.items.as<uint32_t>(), n_items, Wi, kMultiL1Bits, d_hist);
  ExchangePlan plan = exchange(sh, me, d_hist, Wi, me.item_buf, [&](const unsigned long long *d_bases) {
    dev_records_scatter(c, me.items.as<uint32_t>(), n_items, Wi, kMultiL1Bits, nullptr, nullptr, d_bases);
  });
  me.items.release();
  me.scratch.reserve((size_t)(std::max<int64_t>(plan.n_recv, 1) + 16) * Wi * 4);
  dev_sdbg_finish(c, me.item_buf.as<uint32_t>(), me.scratch.as<uint32_t>(), plan.n_recv, plan.chunk_start.data(), plan.chunk_size.data(),
                  plan.chunk_seg.data(), (int)plan.chunk_start.size(), plan.n_segs, k, kMultiL1Bits, tip_mode, &me.sdbg);
  me.scratch.release();
  me.item_buf.release();
}

MultiGpu::MultiGpu(const std::vector<int> &devices) {
  int n = 0;
  MF_CUDA(cudaGetDeviceCount(&n));
  for (size_t i = 0; i < devices.size(); ++i) {
    if (devices[i] < 0 || devices[i] >= n) throw std::invalid_argument("gpu id " + std::to_string(devices[i]) + " is not a visible CUDA device");
    for (size_t j = 0; j < i; ++j)
      if (devices[j] == devices[i]) throw std::invalid_argument("gpu ids must be distinct");
  }
  for (size_t i = 0; i < devices.size(); ++i) {
    ranks.emplace_back(new MultiRank());
    ranks.back()->rank = (int)i;
    ranks.back()->device = devices[i];
  }
  // contexts and peer mappings, once
  for (auto &r : ranks) {
    r->ctx.reset(new Ctx(r->device));
    for (auto &p : ranks) {
      if (p->device == r->device) continue;
      int can = 0;
      MF_CUDA(cudaDeviceCanAccessPeer(&can, r->device, p->device));
      if (!can) throw CudaError("GPUs " + std::to_string(r->device) + " and " + std::to_string(p->device) + " cannot map each other's memory");
      cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cuda_check(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
      cudaGetLastError();
    }
  }
}
MultiGpu::~MultiGpu() {
  for (auto &r : ranks) {
    cudaSetDevice(r->device);
    r.reset();
  }
}
int MultiGpu::world() const { return (int)ranks.size(); }
Ctx &MultiGpu::ctx(int r) { return *ranks[r]->ctx; }
const EdgesView &MultiGpu::edges(int r) const { return ranks[r]->edges; }
const SdbgView &MultiGpu::sdbg(int r) const { return ranks[r]->sdbg; }
const std::vector<int64_t> &MultiGpu::counting(int r) const { return ranks[r]->counting; }

// Runs `body` on one thread per device; the first exception of any rank is rethrown after all threads have joined (a rank that
// fails aborts the barrier, so the others leave their waits instead of deadlocking).
template <class Body>
static void run_ranks(MultiGpu &mg, Body &&body) {
  const int world = mg.world();
  Shared sh;
  sh.world = world;
  sh.bar.reset(new Barrier(world));
  sh.H.assign(world, std::vector<int64_t>(1 << kMultiL1Bits, 0));
  sh.ptrs.assign(world, nullptr);
  sh.err.assign(world, nullptr);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r)
    th.emplace_back([&, r] {
      try {
        MF_CUDA(cudaSetDevice(mg.ranks[r]->device));
        mg.ranks[r]->ctx->begin_call();
        body(sh, *mg.ranks[r]);
        mg.ranks[r]->ctx->end_call();
      } catch (const RankAborted &) {
        // the rank that failed reports
      } catch (...) {
        sh.err[r] = std::current_exception();
        sh.bar->abort();
      }
    });
  for (auto &t : th) t.join();
  for (auto &e : sh.err)
    if (e) std::rethrow_exception(e);
}

void MultiGpu::count(const std::vector<ReadsView> &reads, int k, int min_count, bool want_counting) {
  run_ranks(*this, [&](Shared &sh, MultiRank &me) { rank_count(sh, me, reads[me.rank], k, min_count, want_counting); });
}
void MultiGpu::read2sdbg(const std::vector<ReadsView> &reads, int k, int min_count) {
  run_ranks(*this, [&](Shared &sh, MultiRank &me) {
    rank_count(sh, me, reads[me.rank], k, min_count, false);
    rank_sdbg(sh, me, me.edges.edges, me.edges.n_edges, SeqsView{}, k, 1);
  });
}
void MultiGpu::seq2sdbg(const std::vector<const uint32_t *> &edges, const std::vector<int64_t> &n_edges, const std::vector<SeqsView> &seqs,
                        int k) {
  run_ranks(*this, [&](Shared &sh, MultiRank &me) { rank_sdbg(sh, me, edges[me.rank], n_edges[me.rank], seqs[me.rank], k, 0); });
}

}  // namespace mf
