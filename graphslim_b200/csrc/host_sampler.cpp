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
// Both are reproduced literally (same container type; generator state is imported/exported in
// torch's serialised layout by the Python caller).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "../../include/graphslim_b200.h"

namespace {

struct Mt19937 {
  uint32_t* s;
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
};

struct Entry {
  int32_t local;
  int64_t e;
};

}  // namespace

struct gs_sampler {
  int32_t n;
  const int64_t* rowptr;
  const int32_t* col;
  const float* val;
  int32_t nh;
  int32_t fan[8];
  const int32_t* labels = nullptr;
  std::vector<int32_t> pos;  // node -> local index within the current class, -1 when unseen
};

extern "C" {

gs_sampler* gs_sampler_create(int32_t n_nodes, const int64_t* rowptr, const int32_t* col, const float* val,
                              int32_t n_hops, const int32_t* fanout) {
  if (n_hops < 1 || n_hops > 5) return nullptr;
  gs_sampler* s = new gs_sampler();
  s->n = n_nodes;
  s->rowptr = rowptr;
  s->col = col;
  s->val = val;
  s->nh = n_hops;
  for (int i = 0; i < n_hops; ++i) s->fan[i] = fanout[i];
  s->pos.assign((size_t)n_nodes, -1);
  return s;
}

void gs_sampler_destroy(gs_sampler* s) { delete s; }

void gs_sampler_set_labels(gs_sampler* s, const int32_t* labels) {
  if (s) s->labels = labels;
}

static inline int64_t align16(int64_t x) { return (x + 15) & ~int64_t(15); }

int64_t gs_sampler_sample_step(gs_sampler* S, int32_t n_class, const int64_t* batch, const int64_t* batch_off,
                               const uint8_t* materialise, uint32_t* mt_state, int32_t* mt_left, int32_t* mt_next,
                               uint8_t* out, int64_t out_cap, int64_t* desc) {
  if (!S || !batch || !batch_off || !mt_state || !out || !desc) return GS_EINVAL;
  Mt19937 rng{mt_state, *mt_left, *mt_next};
  const int nh = S->nh;
  const int64_t* rp = S->rowptr;
  const int32_t* gc = S->col;
  const float* gv = S->val;

  // batched outputs
  std::vector<std::vector<int32_t>> seg(nh + 1, std::vector<int32_t>(n_class + 1, 0));
  std::vector<int32_t> nid_all, tcls, target_ids;
  std::vector<float> inv_b;
  struct Blk {
    std::vector<int32_t> rowptr{0}, col, gcol;
    std::vector<float> val;
  };
  std::vector<Blk> blk(nh);

  std::vector<int32_t> nid;  // node list of the current class (level by level, prefix-stable)
  std::vector<Entry> rowbuf;
  for (int32_t c = 0; c < n_class; ++c) {
    const int64_t b0 = batch_off[c], b1 = batch_off[c + 1];
    const bool keep = materialise == nullptr || materialise[c] != 0;
    nid.clear();
    for (int64_t t = b0; t < b1; ++t) {
      const int32_t v = (int32_t)batch[t];
      if (v < 0 || v >= S->n) return GS_EINVAL;
      // duplicates cannot occur in a permutation slice; the reference's map would alias them
      S->pos[v] = (int32_t)nid.size();
      nid.push_back(v);
    }
    std::vector<int32_t> level_count(nh + 1, 0);
    level_count[0] = (int32_t)nid.size();
    for (int h = 0; h < nh; ++h) {
      const int32_t k = S->fan[h];
      const int32_t n_rows = (int32_t)nid.size();
      const bool last = (h == nh - 1);
      if (!keep && last) {
        // only advance the generator: k draws for every row with more than k neighbours
        int64_t draws = 0;
        for (int32_t t = 0; t < n_rows; ++t) {
          const int64_t deg = rp[nid[t] + 1] - rp[nid[t]];
          if (deg > k) draws += k;
        }
        for (int64_t i = 0; i < draws; ++i) (void)rng();
        break;
      }
      const int32_t col_base = keep ? seg[h + 1][c] : 0;
      for (int32_t t = 0; t < n_rows; ++t) {
        const int32_t v = nid[t];
        const int64_t beg = rp[v], deg = rp[v + 1] - beg;
        std::unordered_set<int64_t> chosen;
        if (deg <= k) {
          for (int64_t j = 0; j < deg; ++j) chosen.insert(j);
        } else {
          for (int64_t j = deg - k; j < deg; ++j) {
            const int64_t r = (int64_t)(rng() % (uint32_t)j);
            if (!chosen.insert(r).second) chosen.insert(j);
          }
        }
        rowbuf.clear();
        for (const int64_t& p : chosen) {
          const int64_t e = beg + p;
          const int32_t u = gc[e];
          int32_t loc = S->pos[u];
          if (loc < 0) {
            loc = (int32_t)nid.size();
            S->pos[u] = loc;
            nid.push_back(u);
          }
          rowbuf.push_back(Entry{loc, e});
        }
        if (keep) {
          std::sort(rowbuf.begin(), rowbuf.end(), [](const Entry& a, const Entry& b) { return a.local < b.local; });
          Blk& B = blk[h];
          for (const Entry& en : rowbuf) {
            B.col.push_back(col_base + en.local);
            B.val.push_back(gv[en.e]);
            if (last) B.gcol.push_back(gc[en.e]);
          }
          B.rowptr.push_back((int32_t)B.col.size());
        }
      }
      level_count[h + 1] = (int32_t)nid.size();
    }
    for (int32_t v : nid) S->pos[v] = -1;
    if (keep) {
      const int32_t bsz = level_count[0];
      for (int32_t t = 0; t < bsz; ++t) {
        tcls.push_back(c);
        inv_b.push_back(1.0f / (float)bsz);
        target_ids.push_back(nid[t]);
      }
      nid_all.insert(nid_all.end(), nid.begin(), nid.end());
    }
    for (int l = 0; l <= nh; ++l) seg[l][c + 1] = seg[l][c] + (keep ? level_count[l] : 0);
  }
  *mt_left = rng.left;
  *mt_next = rng.next;

  // ---- pack -------------------------------------------------------------------------------
  for (int i = 0; i < 64; ++i) desc[i] = -1;
  int64_t off = 0;
  auto put = [&](const void* src, int64_t bytes) -> int64_t {
    const int64_t at = off;
    if (at + bytes > out_cap) return -1;
    if (bytes) std::memcpy(out + at, src, (size_t)bytes);
    off = align16(at + bytes);
    return at;
  };
  desc[0] = nh;
  desc[1] = seg[0][n_class];
  for (int l = 0; l <= nh; ++l) desc[2 + l] = seg[l][n_class];
  {
    std::vector<int32_t> flat;
    for (int l = 0; l <= nh; ++l) flat.insert(flat.end(), seg[l].begin(), seg[l].end());
    if ((desc[8] = put(flat.data(), (int64_t)flat.size() * 4)) < 0) return GS_ENOSPC;
  }
  if ((desc[9] = put(nid_all.data(), (int64_t)nid_all.size() * 4)) < 0) return GS_ENOSPC;
  if ((desc[10] = put(tcls.data(), (int64_t)tcls.size() * 4)) < 0) return GS_ENOSPC;
  if ((desc[11] = put(inv_b.data(), (int64_t)inv_b.size() * 4)) < 0) return GS_ENOSPC;
  if ((desc[12] = put(target_ids.data(), (int64_t)target_ids.size() * 4)) < 0) return GS_ENOSPC;
  if (S->labels) {
    std::vector<int32_t> tl(target_ids.size());
    for (size_t i = 0; i < tl.size(); ++i) tl[i] = S->labels[target_ids[i]];
    if ((desc[13] = put(tl.data(), (int64_t)tl.size() * 4)) < 0) return GS_ENOSPC;
  }
  std::vector<int32_t> t_rowptr, t_col, cursor;
  std::vector<float> t_val;
  for (int h = 0; h < nh; ++h) {
    Blk& B = blk[h];
    const int32_t n_rows = seg[h][n_class], n_cols = seg[h + 1][n_class];
    const int64_t nnz = (int64_t)B.col.size();
    if ((int32_t)B.rowptr.size() != n_rows + 1) return GS_EINVAL;
    // transposed structure (sample-time CSC): deterministic, rows of A^T sorted by source row
    t_rowptr.assign((size_t)n_cols + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) t_rowptr[(size_t)B.col[e] + 1]++;
    for (int32_t j = 0; j < n_cols; ++j) t_rowptr[j + 1] += t_rowptr[j];
    cursor.assign(t_rowptr.begin(), t_rowptr.end() - 1);
    t_col.resize((size_t)nnz);
    t_val.resize((size_t)nnz);
    for (int32_t r = 0; r < n_rows; ++r)
      for (int32_t e = B.rowptr[r]; e < B.rowptr[r + 1]; ++e) {
        const int32_t w = cursor[B.col[e]]++;
        t_col[w] = r;
        t_val[w] = B.val[e];
      }
    int64_t* d = desc + 16 + 8 * h;
    d[0] = nnz;
    if ((d[1] = put(B.rowptr.data(), (int64_t)B.rowptr.size() * 4)) < 0) return GS_ENOSPC;
    if ((d[2] = put(B.col.data(), nnz * 4)) < 0) return GS_ENOSPC;
    if ((d[3] = put(B.val.data(), nnz * 4)) < 0) return GS_ENOSPC;
    if ((d[4] = put(t_rowptr.data(), (int64_t)t_rowptr.size() * 4)) < 0) return GS_ENOSPC;
    if ((d[5] = put(t_col.data(), nnz * 4)) < 0) return GS_ENOSPC;
    if ((d[6] = put(t_val.data(), nnz * 4)) < 0) return GS_ENOSPC;
    if (h == nh - 1) {
      if ((d[7] = put(B.gcol.data(), nnz * 4)) < 0) return GS_ENOSPC;
    }
  }
  return off;
}

}  // extern "C"
