// Host side of the GCond real-graph path: bit-exact class-batch neighbour sampling.
//
// Replaces TransAndInd.retrieve_class_sampler (graphslim/dataset/loader.py:187-224) and the
// third-party code under it (torch_geometric NeighborSampler.sample -> torch_sparse sample_adj),
// for all classes of one outer step in a single call, emitting the device-ready batched block
// structures (CSR + transposed CSR per hop, int32/fp32) into one packed pinned buffer.
//
// Bit-exactness contract: the choice of neighbours must equal the
// reference's for the same torch CPU generator state.  That requires
//   * one mt19937 word `% j` per Floyd draw (torch::randint(0, j, {1})), drawn in row order;
//   * the iteration order of libstdc++'s std::unordered_set<int64_t>, which fixes the order in
//     which new nodes are discovered (and therefore which draws later rows receive).
// Both are reproduced literally (the real container, fed from a stack arena; generator state is
// imported/exported in torch's serialised layout by the Python caller).
//
// Speed (this is the host-side bottleneck of an epoch once the GPU side is fast):
//   * the generator stream is serial, but the number of draws a hop consumes depends only on the
//     degrees of its rows.  All hops but the last are cheap (<= 256 * fanout rows per class) and run
//     serially; for the last hop the stream is cut into per-class segments (state snapshot + discard),
//     and the classes are then sampled in parallel by a small thread pool;
//   * the sampler keeps its own interleaved (col,val) copy of the CSR and software-prefetches the
//     rowptr / edge cache lines of a block of rows before consuming them: the work is otherwise
//     dominated by ~2 DRAM misses per sampled edge.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory_resource>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/graphslim_b200.h"

namespace {

struct Mt19937 {
  uint32_t s[624];
  int32_t left, next;
  static inline uint32_t tw(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  void reload() {
    uint32_t* p = s;
    for (int j = 624 - 397 + 1; --j; ++p) *p = p[397] ^ tw(p[0], p[1]);
    for (int j = 397; --j; ++p) *p = p[397 - 624] ^ tw(p[0], p[1]);
    *p = p[397 - 624] ^ tw(p[0], s[0]);
    left = 624;
    next = 0;
  }
  inline uint32_t operator()() {
    if (--left == 0) reload();
    uint32_t y = s[next++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    return y ^ (y >> 18);
  }
  void discard(int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
      if (--left == 0) reload();
      ++next;
    }
  }
};

struct ColVal {
  int32_t col;
  float val;
};

struct Entry {
  int32_t local;
  int64_t e;
};

struct HopOut {
  std::vector<int32_t> rowptr;   // local, starts at 0, n_rows + 1 entries
  std::vector<int32_t> col;      // class-local column ids
  std::vector<int32_t> gcol;     // global ids (last hop only)
  std::vector<float> val;
  // transpose of this class's block (sample-time CSC): rows = class-local columns, entries sorted by source row
  std::vector<int32_t> t_rowptr, t_col;
  std::vector<float> t_val;
  void build_transpose(int32_t n_cols) {
    const int32_t n_rows = (int32_t)rowptr.size() - 1;
    const size_t nnz = col.size();
    t_rowptr.assign((size_t)n_cols + 1, 0);
    for (size_t e = 0; e < nnz; ++e) t_rowptr[(size_t)col[e] + 1]++;
    for (int32_t j = 0; j < n_cols; ++j) t_rowptr[(size_t)j + 1] += t_rowptr[(size_t)j];
    std::vector<int32_t> cursor(t_rowptr.begin(), t_rowptr.end() - 1);
    t_col.resize(nnz);
    t_val.resize(nnz);
    for (int32_t r = 0; r < n_rows; ++r)
      for (int32_t e = rowptr[(size_t)r]; e < rowptr[(size_t)r + 1]; ++e) {
        const int32_t w = cursor[(size_t)col[(size_t)e]]++;
        t_col[(size_t)w] = r;
        t_val[(size_t)w] = val[(size_t)e];
      }
  }
};

struct ClassOut {
  std::vector<int32_t> nid;                 // discovered nodes, level by level (prefix-stable)
  std::vector<int32_t> level_count;         // nh + 1
  std::vector<HopOut> hop;                  // nh
};

constexpr int kPrefetchBlock = 64;

}  // namespace

struct gs_sampler {
  int32_t n;
  const int64_t* rowptr;
  std::vector<ColVal> cv;                   // interleaved copy: one cache line per sampled edge instead of two
  int32_t nh;
  int32_t fan[8];
  const int32_t* labels = nullptr;
  std::vector<int32_t> pos_a;               // phase A (serial hops): node -> class-local index, -1 when unseen
  std::vector<std::vector<int32_t>> pos;    // per phase-B worker
  int n_workers = 1;
  int32_t align = 1;                        // every class segment of every level is padded to a multiple of this
};

namespace {

// One hop for one class.  `nid` = all nodes discovered so far (the rows), appended to in place.
// out == nullptr: nothing is recorded (only nid / pos evolve and the generator advances).
inline void sample_hop(const gs_sampler* S, std::vector<int32_t>& pos, Mt19937& rng, std::vector<int32_t>& nid, int32_t k,
                       bool last, HopOut* out) {
  const int64_t* rp = S->rowptr;
  const ColVal* cv = S->cv.data();
  const int32_t n_rows = (int32_t)nid.size();
  std::vector<Entry> rowbuf;
  rowbuf.reserve(32);
  if (out) {
    out->rowptr.clear();
    out->rowptr.reserve((size_t)n_rows + 1);
    out->rowptr.push_back(0);
    out->col.clear();
    out->val.clear();
    out->gcol.clear();
    out->col.reserve((size_t)n_rows * (size_t)k);
    out->val.reserve((size_t)n_rows * (size_t)k);
    if (last) out->gcol.reserve((size_t)n_rows * (size_t)k);
  }
  // block-wise: draw (needs only degrees) and prefetch the chosen edges, then consume them
  int64_t chosen_e[kPrefetchBlock * 16];
  int32_t chosen_n[kPrefetchBlock];
  for (int32_t b0 = 0; b0 < n_rows; b0 += kPrefetchBlock) {
    const int32_t b1 = std::min(n_rows, b0 + kPrefetchBlock);
    for (int32_t t = b0; t < b1; ++t) __builtin_prefetch(rp + nid[(size_t)t]);
    for (int32_t t = b0; t < b1; ++t) {
      const int32_t v = nid[(size_t)t];
      const int64_t beg = rp[v], deg = rp[v + 1] - beg;
      alignas(16) unsigned char arena[2048];
      std::pmr::monotonic_buffer_resource pool(arena, sizeof(arena));
      std::pmr::unordered_set<int64_t> chosen(&pool);
      if (deg <= k) {
        for (int64_t j = 0; j < deg; ++j) chosen.insert(j);
      } else {
        for (int64_t j = deg - k; j < deg; ++j) {
          const int64_t r = (int64_t)(rng() % (uint32_t)j);
          if (!chosen.insert(r).second) chosen.insert(j);
        }
      }
      int32_t cnt = 0;
      int64_t* dst = chosen_e + (size_t)(t - b0) * 16;
      for (const int64_t& p : chosen) {
        const int64_t e = beg + p;
        dst[cnt++] = e;
        __builtin_prefetch(cv + e);
      }
      chosen_n[t - b0] = cnt;
    }
    for (int32_t t = b0; t < b1; ++t) {
      const int64_t* src = chosen_e + (size_t)(t - b0) * 16;
      const int32_t cnt = chosen_n[t - b0];
      rowbuf.clear();
      for (int32_t i = 0; i < cnt; ++i) {
        const int64_t e = src[i];
        const int32_t u = cv[e].col;
        int32_t loc = pos[(size_t)u];
        if (loc < 0) {
          loc = (int32_t)nid.size();
          pos[(size_t)u] = loc;
          nid.push_back(u);
        }
        rowbuf.push_back(Entry{loc, e});
      }
      if (out) {
        std::sort(rowbuf.begin(), rowbuf.end(), [](const Entry& a, const Entry& b) { return a.local < b.local; });
        for (const Entry& en : rowbuf) {
          out->col.push_back(en.local);
          out->val.push_back(cv[en.e].val);
          if (last) out->gcol.push_back(cv[en.e].col);
        }
        out->rowptr.push_back((int32_t)out->col.size());
      }
    }
  }
}

inline int64_t count_draws(const gs_sampler* S, const std::vector<int32_t>& nid, int32_t k) {
  const int64_t* rp = S->rowptr;
  int64_t draws = 0;
  const int32_t n_rows = (int32_t)nid.size();
  for (int32_t b0 = 0; b0 < n_rows; b0 += 256) {
    const int32_t b1 = std::min(n_rows, b0 + 256);
    for (int32_t t = b0; t < b1; ++t) __builtin_prefetch(rp + nid[(size_t)t]);
    for (int32_t t = b0; t < b1; ++t) {
      const int64_t deg = rp[nid[(size_t)t] + 1] - rp[nid[(size_t)t]];
      if (deg > k) draws += k;
    }
  }
  return draws;
}

}  // namespace

extern "C" {

gs_sampler* gs_sampler_create(int32_t n_nodes, const int64_t* rowptr, const int32_t* col, const float* val,
                              int32_t n_hops, const int32_t* fanout) {
  if (n_hops < 1 || n_hops > 5 || !rowptr || !col || !val) return nullptr;
  for (int i = 0; i < n_hops; ++i)
    if (fanout[i] < 0 || fanout[i] > 15) return nullptr;  // arena / block buffers are sized for the reference fan-outs
  gs_sampler* s = new gs_sampler();
  s->n = n_nodes;
  s->rowptr = rowptr;
  const int64_t nnz = rowptr[n_nodes];
  s->cv.resize((size_t)nnz);
  for (int64_t e = 0; e < nnz; ++e) s->cv[(size_t)e] = ColVal{col[e], val[e]};
  s->nh = n_hops;
  for (int i = 0; i < n_hops; ++i) s->fan[i] = fanout[i];
  unsigned hw = std::thread::hardware_concurrency();
  s->n_workers = (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
  s->pos.assign((size_t)s->n_workers, std::vector<int32_t>());
  s->pos_a.assign((size_t)n_nodes, -1);
  return s;
}

void gs_sampler_destroy(gs_sampler* s) { delete s; }

void gs_sampler_set_labels(gs_sampler* s, const int32_t* labels) {
  if (s) s->labels = labels;
}

void gs_sampler_set_align(gs_sampler* s, int32_t align) {
  if (s && align >= 1) s->align = align;
}

void gs_sampler_set_threads(gs_sampler* s, int32_t n) {
  if (s && n >= 1) {
    s->n_workers = n > 64 ? 64 : n;
    if ((int)s->pos.size() < s->n_workers) s->pos.resize((size_t)s->n_workers);
  }
}

static inline int64_t align16(int64_t x) { return (x + 15) & ~int64_t(15); }

struct gs_sample_job {
  int32_t n_class = 0;
  std::vector<ClassOut> co;
  std::vector<Mt19937> snap;          // generator state at the start of each class's last hop
  std::vector<uint8_t> keepv;
  int64_t us_a = 0;
};

// Phase A (serial, owns the generator): class batches -> all hops but the last, and the per-class segmentation of the
// last hop's random stream.  Independent of phase B of earlier steps, so the two can run on different threads.
gs_sample_job* gs_sampler_begin_step(gs_sampler* S, int32_t n_class, const int64_t* batch, const int64_t* batch_off,
                                     const uint8_t* materialise, uint32_t* mt_state, int32_t* mt_left,
                                     int32_t* mt_next) {
  if (!S || !batch || !batch_off || !mt_state || n_class < 1) return nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  const int nh = S->nh;
  Mt19937 rng;
  std::memcpy(rng.s, mt_state, sizeof(rng.s));
  rng.left = *mt_left;
  rng.next = *mt_next;
  gs_sample_job* J = new gs_sample_job();
  J->n_class = n_class;
  J->co.resize((size_t)n_class);
  J->snap.resize((size_t)n_class);
  J->keepv.resize((size_t)n_class);
  std::vector<ClassOut>& co = J->co;
  std::vector<Mt19937>& snap = J->snap;
  std::vector<uint8_t>& keepv = J->keepv;
  {
    std::vector<int32_t>& pos = S->pos_a;
    for (int32_t c = 0; c < n_class; ++c) {
      ClassOut& C = co[(size_t)c];
      const bool keep = materialise == nullptr || materialise[c] != 0;
      keepv[(size_t)c] = keep;
      C.level_count.assign((size_t)nh + 1, 0);
      C.hop.resize((size_t)nh);
      C.nid.clear();
      for (int64_t t = batch_off[c]; t < batch_off[c + 1]; ++t) {
        const int32_t v = (int32_t)batch[t];
        if (v < 0 || v >= S->n) {
          for (int32_t u : C.nid) pos[(size_t)u] = -1;
          delete J;
          return nullptr;
        }
        pos[(size_t)v] = (int32_t)C.nid.size();
        C.nid.push_back(v);
      }
      C.level_count[0] = (int32_t)C.nid.size();
      for (int h = 0; h + 1 < nh; ++h) {
        sample_hop(S, pos, rng, C.nid, S->fan[h], false, keep ? &C.hop[(size_t)h] : nullptr);
        C.level_count[(size_t)h + 1] = (int32_t)C.nid.size();
      }
      for (int32_t v : C.nid) pos[(size_t)v] = -1;
      snap[(size_t)c] = rng;
      rng.discard(count_draws(S, C.nid, S->fan[nh - 1]));
    }
  }
  std::memcpy(mt_state, rng.s, sizeof(rng.s));
  *mt_left = rng.left;
  *mt_next = rng.next;
  J->us_a = (int64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_start).count();
  return J;
}

// Phase B (parallel over classes) + packing.  Consumes and frees the job.
int64_t gs_sampler_finish_step(gs_sampler* S, gs_sample_job* J, uint8_t* out, int64_t out_cap, int64_t* desc) {
  if (!S || !J || !out || !desc) {
    delete J;
    return GS_EINVAL;
  }
  struct JobGuard {
    gs_sample_job* j;
    ~JobGuard() { delete j; }
  } guard{J};
  const int nh = S->nh;
  const int32_t n_class = J->n_class;
  std::vector<ClassOut>& co = J->co;
  std::vector<Mt19937>& snap = J->snap;
  std::vector<uint8_t>& keepv = J->keepv;
  auto us_since = [&](std::chrono::steady_clock::time_point t0) {
    return (int64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
  };
  const int64_t us_a = J->us_a;
  const auto t_b = std::chrono::steady_clock::now();
  // ---- phase B (parallel over classes): the last hop, each class on its own segment of the stream
  {
    std::atomic<int32_t> next_class{0};
    auto work = [&](int w) {
      std::vector<int32_t>& pos = S->pos[(size_t)w];
      if ((int32_t)pos.size() != S->n) pos.assign((size_t)S->n, -1);
      for (;;) {
        const int32_t c = next_class.fetch_add(1);
        if (c >= n_class) break;
        if (!keepv[(size_t)c]) continue;
        ClassOut& C = co[(size_t)c];
        for (size_t i = 0; i < C.nid.size(); ++i) pos[(size_t)C.nid[i]] = (int32_t)i;
        Mt19937 local = snap[(size_t)c];
        sample_hop(S, pos, local, C.nid, S->fan[nh - 1], true, &C.hop[(size_t)nh - 1]);
        C.level_count[(size_t)nh] = (int32_t)C.nid.size();
        for (int32_t v : C.nid) pos[(size_t)v] = -1;
        for (int h = 0; h < nh; ++h) C.hop[(size_t)h].build_transpose(C.level_count[(size_t)h + 1]);
      }
    };
    int n_keep = 0;
    for (int32_t c = 0; c < n_class; ++c) n_keep += keepv[(size_t)c];
    const int nw = std::max(1, std::min(S->n_workers, n_keep));
    if (nw == 1) {
      work(0);
    } else {
      std::vector<std::thread> th;
      for (int w = 1; w < nw; ++w) th.emplace_back(work, w);
      work(0);
      for (auto& t : th) t.join();
    }
  }

  const int64_t us_b = us_since(t_b);
  const auto t_c = std::chrono::steady_clock::now();
  // ---- batch offsets
  std::vector<std::vector<int32_t>> seg((size_t)nh + 1, std::vector<int32_t>((size_t)n_class + 1, 0));
  // (padded to S->align rows per class: pad rows carry no edges, zero loss weight and node id 0, so they contribute
  //  nothing, but every class segment then starts on a tile boundary of the tensor-core grouped products)
  const int32_t al = S->align;
  for (int32_t c = 0; c < n_class; ++c)
    for (int l = 0; l <= nh; ++l) {
      const int32_t cnt = keepv[(size_t)c] ? co[(size_t)c].level_count[(size_t)l] : 0;
      seg[(size_t)l][(size_t)c + 1] = seg[(size_t)l][(size_t)c] + (cnt + al - 1) / al * al;
    }

  // ---- pack ---------------------------------------------------------------------------------
  for (int i = 0; i < 64; ++i) desc[i] = -1;
  int64_t off = 0;
  auto reserve = [&](int64_t bytes) -> int64_t {
    const int64_t at = off;
    if (at + bytes > out_cap) return -1;
    off = align16(at + bytes);
    return at;
  };
  desc[0] = nh;
  desc[1] = seg[0][(size_t)n_class];
  for (int l = 0; l <= nh; ++l) desc[2 + l] = seg[(size_t)l][(size_t)n_class];
  const int64_t n_tgt = seg[0][(size_t)n_class], n_last = seg[(size_t)nh][(size_t)n_class];
  {
    const int64_t bytes = (int64_t)(nh + 1) * (n_class + 1) * 4;
    if ((desc[8] = reserve(bytes)) < 0) return GS_ENOSPC;
    int32_t* d = reinterpret_cast<int32_t*>(out + desc[8]);
    for (int l = 0; l <= nh; ++l)
      std::memcpy(d + (size_t)l * (size_t)(n_class + 1), seg[(size_t)l].data(), (size_t)(n_class + 1) * 4);
  }
  if ((desc[9] = reserve(n_last * 4)) < 0) return GS_ENOSPC;
  if ((desc[10] = reserve(n_tgt * 4)) < 0) return GS_ENOSPC;
  if ((desc[11] = reserve(n_tgt * 4)) < 0) return GS_ENOSPC;
  if ((desc[12] = reserve(n_tgt * 4)) < 0) return GS_ENOSPC;
  if (S->labels && (desc[13] = reserve(n_tgt * 4)) < 0) return GS_ENOSPC;
  if ((desc[14] = reserve((int64_t)(nh + 1) * n_class * 4)) < 0) return GS_ENOSPC;   // true (unpadded) counts
  {
    int32_t* cnt = reinterpret_cast<int32_t*>(out + desc[14]);
    for (int l = 0; l <= nh; ++l)
      for (int32_t c = 0; c < n_class; ++c)
        cnt[(size_t)l * (size_t)n_class + (size_t)c] = keepv[(size_t)c] ? co[(size_t)c].level_count[(size_t)l] : 0;
  }
  {
    int32_t* nid_all = reinterpret_cast<int32_t*>(out + desc[9]);
    int32_t* tcls = reinterpret_cast<int32_t*>(out + desc[10]);
    float* inv_b = reinterpret_cast<float*>(out + desc[11]);
    int32_t* tgt = reinterpret_cast<int32_t*>(out + desc[12]);
    int32_t* tlab = S->labels ? reinterpret_cast<int32_t*>(out + desc[13]) : nullptr;
    for (int32_t c = 0; c < n_class; ++c) {
      if (!keepv[(size_t)c]) continue;
      const ClassOut& C = co[(size_t)c];
      int32_t* dst = nid_all + seg[(size_t)nh][(size_t)c];
      std::memcpy(dst, C.nid.data(), C.nid.size() * 4);
      for (int32_t t = (int32_t)C.nid.size(); t < seg[(size_t)nh][(size_t)c + 1] - seg[(size_t)nh][(size_t)c]; ++t)
        dst[t] = 0;
      const int32_t bsz = C.level_count[0];
      const int32_t o = seg[0][(size_t)c];
      for (int32_t t = 0; t < bsz; ++t) {
        tcls[o + t] = c;
        inv_b[o + t] = 1.0f / (float)bsz;
        tgt[o + t] = C.nid[(size_t)t];
        if (tlab) tlab[o + t] = S->labels[C.nid[(size_t)t]];
      }
      for (int32_t t = bsz; t < seg[0][(size_t)c + 1] - o; ++t) {     // pad targets: zero weight in the loss
        tcls[o + t] = c;
        inv_b[o + t] = 0.0f;
        tgt[o + t] = 0;
        if (tlab) tlab[o + t] = 0;
      }
    }
  }
  std::vector<int32_t> cursor;
  for (int h = 0; h < nh; ++h) {
    const int32_t n_rows = seg[(size_t)h][(size_t)n_class], n_cols = seg[(size_t)h + 1][(size_t)n_class];
    int64_t nnz = 0;
    for (int32_t c = 0; c < n_class; ++c)
      if (keepv[(size_t)c]) nnz += (int64_t)co[(size_t)c].hop[(size_t)h].col.size();
    int64_t* d = desc + 16 + 8 * h;
    d[0] = nnz;
    if ((d[1] = reserve((int64_t)(n_rows + 1) * 4)) < 0) return GS_ENOSPC;
    if ((d[2] = reserve(nnz * 4)) < 0) return GS_ENOSPC;
    if ((d[3] = reserve(nnz * 4)) < 0) return GS_ENOSPC;
    if ((d[4] = reserve((int64_t)(n_cols + 1) * 4)) < 0) return GS_ENOSPC;
    if ((d[5] = reserve(nnz * 4)) < 0) return GS_ENOSPC;
    if ((d[6] = reserve(nnz * 4)) < 0) return GS_ENOSPC;
    if (h == nh - 1 && (d[7] = reserve(nnz * 4)) < 0) return GS_ENOSPC;
    int32_t* rowptr = reinterpret_cast<int32_t*>(out + d[1]);
    int32_t* col = reinterpret_cast<int32_t*>(out + d[2]);
    float* val = reinterpret_cast<float*>(out + d[3]);
    int32_t* t_rowptr = reinterpret_cast<int32_t*>(out + d[4]);
    int32_t* t_col = reinterpret_cast<int32_t*>(out + d[5]);
    float* t_val = reinterpret_cast<float*>(out + d[6]);
    int32_t* gcol = (h == nh - 1) ? reinterpret_cast<int32_t*>(out + d[7]) : nullptr;
    // forward structure: concatenate classes, shifting local columns by the class offset of the next level
    int64_t e0 = 0;
    rowptr[0] = 0;
    for (int32_t c = 0; c < n_class; ++c) {
      if (!keepv[(size_t)c]) continue;
      const HopOut& H = co[(size_t)c].hop[(size_t)h];
      const int32_t r0 = seg[(size_t)h][(size_t)c], cbase = seg[(size_t)h + 1][(size_t)c];
      const int32_t rows_c = (int32_t)H.rowptr.size() - 1;
      if (rows_c != co[(size_t)c].level_count[(size_t)h]) return GS_EINVAL;
      for (int32_t r = 0; r < rows_c; ++r) rowptr[r0 + r + 1] = (int32_t)(e0 + H.rowptr[(size_t)r + 1]);
      const size_t m = H.col.size();
      for (int32_t r = r0 + rows_c; r < seg[(size_t)h][(size_t)c + 1]; ++r) rowptr[r + 1] = (int32_t)(e0 + (int64_t)m);
      for (size_t i = 0; i < m; ++i) col[e0 + (int64_t)i] = H.col[i] + cbase;
      if (m) std::memcpy(val + e0, H.val.data(), m * 4);
      if (gcol && m) std::memcpy(gcol + e0, H.gcol.data(), m * 4);
      e0 += (int64_t)m;
    }
    // transposed structure: the batch is block diagonal, so it is the concatenation of the per-class transposes
    // (built in parallel above) with row / entry offsets added
    {
      int64_t te = 0;
      t_rowptr[0] = 0;
      for (int32_t c = 0; c < n_class; ++c) {
        const int32_t c0 = seg[(size_t)h + 1][(size_t)c], c1 = seg[(size_t)h + 1][(size_t)c + 1];
        if (!keepv[(size_t)c]) continue;
        const HopOut& H = co[(size_t)c].hop[(size_t)h];
        const int32_t cols_c = (int32_t)H.t_rowptr.size() - 1;
        const int32_t rbase = seg[(size_t)h][(size_t)c];
        for (int32_t j = 0; j < cols_c; ++j) t_rowptr[c0 + j + 1] = (int32_t)(te + H.t_rowptr[(size_t)j + 1]);
        const size_t m = H.t_col.size();
        for (size_t i = 0; i < m; ++i) t_col[te + (int64_t)i] = H.t_col[i] + rbase;
        if (m) std::memcpy(t_val + te, H.t_val.data(), m * 4);
        te += (int64_t)m;
        for (int32_t j = c0 + cols_c; j < c1; ++j) t_rowptr[j + 1] = (int32_t)te;      // pad columns: empty rows
      }
    }
  }
  desc[40] = us_a;                // serial hops + stream segmentation
  desc[41] = us_b;                // parallel last hop
  desc[42] = us_since(t_c);       // packing (offsets, transposes)
  return off;
}


int64_t gs_sampler_sample_step(gs_sampler* S, int32_t n_class, const int64_t* batch, const int64_t* batch_off,
                               const uint8_t* materialise, uint32_t* mt_state, int32_t* mt_left, int32_t* mt_next,
                               uint8_t* out, int64_t out_cap, int64_t* desc) {
  if (!S || !batch || !batch_off || !mt_state || !out || !desc || n_class < 1) return GS_EINVAL;
  gs_sample_job* J = gs_sampler_begin_step(S, n_class, batch, batch_off, materialise, mt_state, mt_left, mt_next);
  if (!J) return GS_EINVAL;
  return gs_sampler_finish_step(S, J, out, out_cap, desc);
}


// ---------------------------------------------------------------------------------------------------------------
// Class batches: `np.random.permutation(members_of_class)[:batch]` for every class in order
// (graphslim/dataset/loader.py:222, numpy's GLOBAL legacy generator).  numpy's legacy RandomState.permutation of a 1-D
// array copies it and runs `_shuffle_raw`: for i = n-1 .. 1: j = random_interval(i); swap(x[i], x[j]), where
// random_interval(max) masks 32-bit MT19937 outputs with the smallest 2^k - 1 >= max and rejects values > max (64-bit
// outputs only for max > 0xffffffff, which cannot occur here).  The stream is frozen ("legacy"), so restating it keeps
// the index selection bit-exact -- and takes the shuffles (154 k elements per outer step at the Reddit shape, 2 ms in
// numpy on the box's host) off the interpreter lock: the call runs without the GIL next to the thread that issues the
// kernels.  `key` (624 words) and `pos` are np.random.get_state()[1:3], updated in place for np.random.set_state().
static inline void np_mt19937_gen(uint32_t* mt) {
  constexpr int N = 624, M = 397;
  constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
  int i = 0;
  uint32_t y;
  for (; i < N - M; ++i) {
    y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
    mt[i] = mt[i + M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  for (; i < N - 1; ++i) {
    y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
    mt[i] = mt[i + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
  mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}

// tempered outputs of the current state block, consumed sequentially (the tempering vectorises; a draw is then a load)
struct NpStream {
  uint32_t* mt;
  uint32_t out[624];
  int pos;
  static inline uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  NpStream(uint32_t* key, int p) : mt(key), pos(p) {
    for (int i = 0; i < 624; ++i) out[i] = temper(mt[i]);
  }
  inline uint32_t next32() {
    if (pos >= 624) {
      np_mt19937_gen(mt);
      for (int i = 0; i < 624; ++i) out[i] = temper(mt[i]);
      pos = 0;
    }
    return out[pos++];
  }
};

int gs_np_legacy_class_batches(uint32_t* key, int32_t* pos_io, int32_t n_class, const int64_t* members,
                               const int64_t* member_off, int32_t batch, int32_t* out, int32_t* out_off) {
  if (!key || !pos_io || n_class < 0 || !member_off || !out || !out_off || batch < 0 || *pos_io < 0 || *pos_io > 624)
    return GS_EINVAL;
  NpStream rng(key, *pos_io);
  std::vector<int64_t> x;
  int32_t w = 0;
  out_off[0] = 0;
  for (int c = 0; c < n_class; ++c) {
    const int64_t n = member_off[c + 1] - member_off[c];
    if (n < 0 || (n > 0 && !members) || n > 0xffffffffLL) return GS_EINVAL;
    x.assign(members + member_off[c], members + member_off[c] + n);
    int64_t* xp = x.data();
    uint64_t mask = 0;                                   // smallest 2^k - 1 >= i, kept up to date as i falls
    if (n > 1) {
      mask = (uint64_t)(n - 1);
      mask |= mask >> 1;
      mask |= mask >> 2;
      mask |= mask >> 4;
      mask |= mask >> 8;
      mask |= mask >> 16;
      mask |= mask >> 32;
    }
    for (int64_t i = n - 1; i >= 1; --i) {
      if ((uint64_t)i <= (mask >> 1)) mask >>= 1;
      uint64_t j;
      do {
        j = (uint64_t)rng.next32() & mask;
      } while (j > (uint64_t)i);
      const int64_t t = xp[i];
      xp[i] = xp[j];
      xp[j] = t;
    }
    const int64_t take = n < batch ? n : batch;
    for (int64_t i = 0; i < take; ++i) {
      if (xp[i] < 0 || xp[i] > 0x7fffffffLL) return GS_EINVAL;
      out[w++] = (int32_t)xp[i];
    }
    out_off[c + 1] = w;
  }
  *pos_io = rng.pos;
  return GS_OK;
}

}  // extern "C"
