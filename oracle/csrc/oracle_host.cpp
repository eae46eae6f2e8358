// ORACLE (test infrastructure, not product code).
//
// CPU restatement of the third-party native pieces the reference's GCond path
// executes but that are not vendored under /root/reference:
//
//   * torch_sparse==0.6.18 `sample_adj` (CPU), reached from
//     graphslim/dataset/loader.py:216-223 through torch_geometric's
//     NeighborSampler.sample -> SparseTensor.sample_adj.  Restated from the
//     published algorithm (csrc/cpu/sample_cpu.cpp upstream): targets are
//     seeded into n_id first; per row, all neighbours are taken when
//     deg <= k, otherwise Robert Floyd's sampling without replacement draws
//     `uniform_randint(j)` for j = deg-k .. deg-1 into a
//     std::unordered_set<int64_t>; the set is iterated, unseen columns are
//     appended to n_id in iteration order; every output row is sorted by its
//     relabelled column.  PARITY UNPINNED at this boundary: torch_sparse is
//     absent from the container so the restatement cannot be executed against
//     the wheel.
//   * torch's CPU generator: `torch::randint(0, j, {1})` is one 32-bit
//     mt19937 output modulo j (verified against torch in
//     tests/test_oracle_rng.py).
//   * torch_sparse `spmm` CPU (csrc/cpu/spmm_cpu.cpp upstream): row-wise
//     sequential accumulation in CSR order, fp32.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// Standard MT19937 with the (left, next) bookkeeping torch's CPU generator
// serialises in get_rng_state().
struct Mt {
  uint32_t* s;  // 624 words
  int32_t left;
  int32_t next;
  static inline uint32_t mix(uint32_t u, uint32_t v) {
    uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  void refill() {
    const int N = 624, M = 397;
    for (int i = 0; i < N - M; ++i) s[i] = s[i + M] ^ mix(s[i], s[i + 1]);
    for (int i = N - M; i < N - 1; ++i) s[i] = s[i + M - N] ^ mix(s[i], s[i + 1]);
    s[N - 1] = s[M - 1] ^ mix(s[N - 1], s[0]);
    left = N;
    next = 0;
  }
  uint32_t draw() {
    if (--left == 0) refill();
    uint32_t y = s[next++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
};

}  // namespace

extern "C" {

// Draw `n` values `mt() % high[i]` -- used to pin the generator restatement
// against torch.randint.
void oracle_mt_randint(uint32_t* state, int32_t* left, int32_t* next, const int64_t* high,
                       int64_t n, int64_t* out) {
  Mt g{state, *left, *next};
  for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)(g.draw() % (uint32_t)high[i]);
  *left = g.left;
  *next = g.next;
}

// One hop of neighbour sampling without replacement.
// Outputs: out_rowptr[n_idx+1]; out_col/out_eid (capacity n_idx*k);
// out_nid (capacity n_idx + n_idx*k).  Returns number of nodes in out_nid.
int64_t oracle_sample_adj(const int64_t* rowptr, const int64_t* col, const int64_t* idx,
                          int64_t n_idx, int64_t k, uint32_t* state, int32_t* left,
                          int32_t* next, int64_t* out_rowptr, int64_t* out_col,
                          int64_t* out_eid, int64_t* out_nid) {
  Mt g{state, *left, *next};
  std::vector<std::vector<std::pair<int64_t, int64_t>>> rows((size_t)n_idx);
  std::vector<int64_t> nid;
  std::unordered_map<int64_t, int64_t> pos;
  nid.reserve((size_t)(n_idx * (k + 1)));
  for (int64_t t = 0; t < n_idx; ++t) {
    pos[idx[t]] = t;
    nid.push_back(idx[t]);
  }
  out_rowptr[0] = 0;
  for (int64_t t = 0; t < n_idx; ++t) {
    const int64_t v = idx[t];
    const int64_t beg = rowptr[v], deg = rowptr[v + 1] - rowptr[v];
    std::unordered_set<int64_t> chosen;
    if (deg <= k) {
      for (int64_t j = 0; j < deg; ++j) chosen.insert(j);
    } else {
      for (int64_t j = deg - k; j < deg; ++j) {
        const int64_t r = (int64_t)(g.draw() % (uint32_t)j);
        if (!chosen.insert(r).second) chosen.insert(j);
      }
    }
    for (const int64_t& p : chosen) {
      const int64_t e = beg + p;
      const int64_t c = col[e];
      auto it = pos.find(c);
      int64_t local;
      if (it == pos.end()) {
        local = (int64_t)nid.size();
        pos[c] = local;
        nid.push_back(c);
      } else {
        local = it->second;
      }
      rows[(size_t)t].emplace_back(local, e);
    }
    out_rowptr[t + 1] = out_rowptr[t] + (int64_t)rows[(size_t)t].size();
  }
  int64_t w = 0;
  for (auto& r : rows) {
    std::sort(r.begin(), r.end(),
              [](const std::pair<int64_t, int64_t>& a, const std::pair<int64_t, int64_t>& b) {
                return a.first < b.first;
              });
    for (auto& pr : r) {
      out_col[w] = pr.first;
      out_eid[w] = pr.second;
      ++w;
    }
  }
  std::memcpy(out_nid, nid.data(), nid.size() * sizeof(int64_t));
  *left = g.left;
  *next = g.next;
  return (int64_t)nid.size();
}

// Y = A @ X, CSR, fp32, sequential accumulation per row in storage order.
void oracle_spmm_csr_f32(int64_t n_rows, const int64_t* rowptr, const int64_t* col,
                         const float* val, const float* X, int64_t F, float* Y) {
  for (int64_t r = 0; r < n_rows; ++r) {
    float* y = Y + r * F;
    for (int64_t f = 0; f < F; ++f) y[f] = 0.f;
    for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const float a = val[e];
      const float* x = X + col[e] * F;
      for (int64_t f = 0; f < F; ++f) y[f] += a * x[f];
    }
  }
}

// Iteration order of the real std::unordered_set<int64_t> after inserting keys[0..n) (duplicates ignored):
// pins the container restatement the device sampler uses.  Returns the number of distinct keys.
int64_t oracle_uset_order(const int64_t* keys, int64_t n, int64_t* out) {
  std::unordered_set<int64_t> s;
  for (int64_t i = 0; i < n; ++i) s.insert(keys[i]);
  int64_t m = 0;
  for (const int64_t& k : s) out[m++] = k;
  return m;
}

}  // extern "C"
