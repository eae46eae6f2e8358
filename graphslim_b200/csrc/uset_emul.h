// Iteration order of libstdc++'s std::unordered_set<int64_t> for at most 16 insertions, without the container.
//
// torch_sparse's sample_adj (csrc/cpu/sample_cpu.cpp, called through torch_geometric's NeighborSampler at
// graphslim/dataset/loader.py:216-223) collects the Floyd-sampled neighbour positions of a row in a
// std::unordered_set<int64_t> and then *iterates* it: the iteration order fixes the order in which new nodes are
// appended to n_id, i.e. which random draws every later row receives.  Bit-exact index selection therefore needs that
// order.  The host sampler feeds the real container; the device sampler (one thread per row) uses this restatement of
// the container's algorithm (libstdc++ hashtable.h / hashtable_policy.h):
//   * std::hash<int64_t> is the identity, bucket = key % bucket_count;
//   * all nodes live on one singly linked list; a bucket stores the node *before* its first node;
//     inserting into an empty bucket puts the node at the list head, otherwise right after the bucket's before-node
//     (_M_insert_bucket_begin);
//   * _Prime_rehash_policy with max_load_factor 1: 1 bucket when empty, 13 at the first insertion, 29 at the 14th
//     (_M_need_rehash / _M_next_bkt), nodes re-linked in list order by _M_rehash_aux(unique keys).
// tests/test_uset_emul.py checks it against the real container (oracle/csrc/oracle_host.cpp) on random key sequences.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD inline
#endif

namespace gs {

struct USetEmul {
  static constexpr int kMax = 16;          // capacity (fan-outs of the reference are <= 15)
  static constexpr int kNil = -1;          // null next pointer
  static constexpr int kEmpty = -2;        // bucket without nodes
  static constexpr int kBeforeBegin = -3;  // the list's before-begin sentinel as a bucket's before-node
  int32_t key[kMax];
  int8_t nxt[kMax];
  int8_t bkt[40];
  int8_t head;
  int32_t nb, cnt, next_resize;

  GS_HD void clear() {
    nb = 1;
    cnt = 0;
    next_resize = 0;
    head = kNil;
    bkt[0] = kEmpty;
  }
  GS_HD int next_of(int node) const { return node == kBeforeBegin ? head : nxt[node]; }
  GS_HD void set_next(int node, int v) {
    if (node == kBeforeBegin) head = (int8_t)v;
    else nxt[node] = (int8_t)v;
  }
  // _Prime_rehash_policy::_M_next_bkt for the sizes reachable with <= 16 elements
  GS_HD int next_bkt(int n) {
    int r;
    if (n <= 13) {
      r = n <= 2 ? 2 : (n == 3 ? 3 : (n <= 5 ? 5 : (n <= 7 ? 7 : (n <= 11 ? 11 : 13))));
      if (n == 0) r = 1;
    } else {
      r = n <= 17 ? 17 : (n <= 19 ? 19 : (n <= 23 ? 23 : (n <= 29 ? 29 : (n <= 31 ? 31 : 37))));
    }
    next_resize = r;
    return r;
  }
  // _M_rehash_aux(n, true_type)
  GS_HD void rehash(int n) {
    int8_t nb2[40];
    for (int i = 0; i < n; ++i) nb2[i] = kEmpty;
    int p = head;
    head = kNil;
    int bbegin_bkt = 0;
    while (p != kNil) {
      const int nx = nxt[p];
      const int b = key[p] % n;
      if (nb2[b] == kEmpty) {
        nxt[p] = head;
        head = (int8_t)p;
        nb2[b] = kBeforeBegin;
        if (nxt[p] != kNil) nb2[bbegin_bkt] = (int8_t)p;
        bbegin_bkt = b;
      } else {
        const int prev = nb2[b];
        nxt[p] = (int8_t)next_of(prev);
        set_next(prev, p);
      }
      p = nx;
    }
    for (int i = 0; i < n; ++i) bkt[i] = nb2[i];
    nb = n;
  }
  // insert(k).second
  GS_HD bool insert(int32_t k) {
    for (int i = 0; i < cnt; ++i)
      if (key[i] == k) return false;
    if (cnt + 1 > next_resize) {                              // _M_need_rehash(nb, cnt, 1)
      int min_bkts = cnt + 1;
      if (next_resize == 0 && min_bkts < 11) min_bkts = 11;
      if (min_bkts >= nb) {
        const int want = (min_bkts + 1 > nb * 2) ? min_bkts + 1 : nb * 2;
        rehash(next_bkt(want));
      } else {
        next_resize = nb;
      }
    }
    const int b = k % nb;
    const int node = cnt;
    key[node] = k;
    if (bkt[b] != kEmpty) {                                   // _M_insert_bucket_begin
      const int prev = bkt[b];
      nxt[node] = (int8_t)next_of(prev);
      set_next(prev, node);
    } else {
      nxt[node] = head;
      head = (int8_t)node;
      if (nxt[node] != kNil) bkt[key[nxt[node]] % nb] = (int8_t)node;
      bkt[b] = kBeforeBegin;
    }
    ++cnt;
    return true;
  }
};

}  // namespace gs
