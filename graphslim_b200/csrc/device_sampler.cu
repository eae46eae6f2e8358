// Class-batch neighbour sampling ON THE DEVICE, bit-exact with the host sampler (host_sampler.cpp) and therefore with
// TransAndInd.retrieve_class_sampler -> torch_geometric NeighborSampler -> torch_sparse sample_adj
// (graphslim/dataset/loader.py:187-224), for all classes of one outer step.
//
// Why: on the host the sampler is the bottleneck of an epoch once the GPU side is fast (one serial mt19937 stream,
// ~2 DRAM misses per sampled edge, then a 30+ MB packed H2D copy per outer step at the Reddit shape).  On the
// device the graph is already resident, the blocks never cross PCIe, and the only serial part left is tiny.
//
// What makes the reference's index selection reproducible in parallel:
//   * torch::randint(0, j) on the CPU generator is the next mt19937 word % j.  The raw (untempered) state blocks of
//     the generator are produced ahead of time by one CTA (mt_generate_kernel: 624-word blocks, three dependent
//     phases each), so draw number p of the step is a plain array read (stream_raw);
//   * a row consumes `fanout` draws iff degree > fanout, so the draw offset of every row is a prefix sum over degrees;
//   * the order in which a row's chosen neighbours are appended to n_id is the iteration order of libstdc++'s
//     unordered_set, restated in uset_emul.h (one thread per row, <= 15 keys);
//   * "first occurrence wins" relabelling = atomicMin over candidate ranks + a prefix sum over the first-occurrence
//     flags, which reproduces the sequential append order exactly.
// Serial dependencies that remain: class c+1's stream offset needs class c's hop-1 result (the number of hop-2 rows).
// sample_serial_kernel (ONE CTA) therefore runs all hops but the last for the classes in order (<= 256*fanout rows
// each) and cuts the stream into per-class segments; sample_last_kernel then runs the last hop (the bulk of the work),
// the per-class transposes and the pos-map cleanup with one CTA per class.  pack_* lay the per-class pieces out in the
// batched block-diagonal format of the host sampler (same `desc` table), so the consumers do not change.
#include <algorithm>
#include <climits>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "uset_emul.h"

namespace gs {
namespace ds {

constexpr int kMaxHops = 5;
constexpr int kThreads = 512;             // 128 registers per thread: the per-row arrays stay out of local memory
constexpr int kWarps = kThreads / 32;
constexpr int kCursorSmemInts = 40 * 1024;       // transposed-block cursors kept in shared memory up to this many columns

struct MtDev {
  uint32_t s[624];
  int32_t left, next;
};

struct Geom {
  int32_t n, nh, n_class_max, batch_max, align;
  int32_t fan[kMaxHops];
  int64_t lcap[kMaxHops + 1];   // rows per class at level l (capacity)
  int64_t ncap[kMaxHops];       // sampled edges per class at hop h (capacity)
};

struct Ptrs {
  const int32_t* rowptr;
  const int32_t* col;
  const float* val;
  const int32_t* labels;
  MtDev* mt;
  uint32_t* R;               // raw state blocks: R[0..624) = state at step start, then regenerated blocks
  int64_t* step_info;        // [0] draws consumed by the step, [1] words left in block 0 at step start, [2] next index
  int32_t* pos;              // [class][n]    node -> class-local index, -1 = unseen
  int32_t* firstq;           // [class][n]    smallest candidate rank that proposed the node in the current hop
  int32_t* nid;              // [class][lcap[nh]]
  int32_t* level_count;      // [class][nh+1]
  int64_t* last_off;         // [class]       stream offset of the class's last hop
  int64_t* hop_off;          // [class]       stream offset of the hop before the last one
  long long* hop_tot;        // [class]       (draws << 32 | candidates) of that hop (phase 1), or the last hop's draws when nh == 1
  int32_t* rowoff[kMaxHops];   // [class][lcap[h]+1]  CSR row pointer of the class block
  int32_t* drawoff[kMaxHops];  // [class][lcap[h]]
  int32_t* cand_e[kMaxHops];   // [class][ncap[h]]    sampled edge ids in discovery order
  int32_t* cand_u[kMaxHops];   // [class][ncap[h]]    their global column ids (kept as gcol after sorting)
  int32_t* ocol[kMaxHops];     // [class][ncap[h]]    class-local columns, rows sorted
  float* oval[kMaxHops];
  int32_t* erow[kMaxHops];     // [class][ncap[h]]    source row of each entry
  int32_t* t_rowptr[kMaxHops]; // [class][lcap[h+1]+1]
  int32_t* t_col[kMaxHops];
  float* t_val[kMaxHops];
  int32_t* t_cursor;           // [class][lcap[nh]]   fallback cursors when the columns do not fit in shared memory
  int32_t* seg;                // [(nh+1)][n_class+1] padded level offsets of the batch
  int64_t* eoff;               // [nh][n_class+1]     edge offsets of the batch
  long long* dbg;              // [16] phase cycle counters of the serial kernel (GS_DS_PROFILE builds)
};

// ------------------------------------------------------------------------------------------ mt19937 stream
__device__ __forceinline__ uint32_t mt_tw(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

// R[0] <- current state; R[b] <- regeneration of R[b-1].  One CTA; the three phases of a block only depend on the
// previous phase (new[i] = new[i-227] ^ tw(old[i], old[i+1]) for i >= 227), so a block costs three barriers.
__global__ void __launch_bounds__(256) mt_generate_kernel(const MtDev* __restrict__ mt, uint32_t* __restrict__ R,
                                                          int nblocks, int64_t* __restrict__ step_info) {
  __shared__ uint32_t cur[2][624];
  const int tid = threadIdx.x;
  for (int i = tid; i < 624; i += 256) {
    const uint32_t v = mt->s[i];
    cur[0][i] = v;
    R[i] = v;
  }
  if (tid == 0) {
    step_info[1] = (int64_t)mt->left - 1;     // reads still available in the current block
    step_info[2] = (int64_t)mt->next;
  }
  __syncthreads();
  for (int b = 1; b <= nblocks; ++b) {
    const uint32_t* o = cur[(b - 1) & 1];
    uint32_t* nw = cur[b & 1];
    uint32_t* dst = R + (size_t)b * 624;
    if (tid < 227) {
      const uint32_t v = o[tid + 397] ^ mt_tw(o[tid], o[tid + 1]);
      nw[tid] = v;
      dst[tid] = v;
    }
    __syncthreads();
    if (tid < 227) {
      const int i = tid + 227;
      const uint32_t v = nw[i - 227] ^ mt_tw(o[i], o[i + 1]);
      nw[i] = v;
      dst[i] = v;
    }
    __syncthreads();
    if (tid < 169) {
      const int i = tid + 454;
      const uint32_t v = nw[i - 227] ^ mt_tw(o[i], o[i + 1]);
      nw[i] = v;
      dst[i] = v;
    } else if (tid == 169) {
      // the last word needs nw[396] (second phase) and nw[0] (first phase): it rides along with the third phase
      const uint32_t v = nw[396] ^ mt_tw(o[623], nw[0]);
      nw[623] = v;
      dst[623] = v;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ block-wide scans
// exclusive prefix sums of f(i), i in [0, m), in index order; emit(i, prefix, f(i)); returns the total to every thread.
// sm: kWarps + 1 elements of shared memory.
template <typename T, class F, class E>
__device__ __forceinline__ T block_scan(int m, F f, E emit, T* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  T carry = 0;
  for (int base = 0; base < m; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const T v = (i < m) ? f(i) : (T)0;
    T x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const T t = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += t;
    }
    if (lane == 31) sm[warp] = x;
    __syncthreads();
    if (warp == 0) {
      T w = (lane < nwarps) ? sm[lane] : (T)0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const T t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      if (lane < nwarps) sm[lane] = w;
    }
    __syncthreads();
    const T wpre = warp > 0 ? sm[warp - 1] : (T)0;
    const T chunk_total = sm[nwarps - 1];
    if (i < m) emit(i, carry + wpre + x - v, v);
    carry += chunk_total;
    __syncthreads();
  }
  return carry;
}

// ------------------------------------------------------------------------------------------ one hop of one class
// Rows = the class's nodes discovered so far (nid[0..n_rows)).  Returns the new node count; *draws = words consumed.
// Every per-row step first issues all of its independent loads (stream words, column ids, pos entries, values) into
// register arrays and only then consumes them: the kernel is a chain of dependent global round trips, so the number of
// round trips per hop, not bandwidth, sets its duration.
constexpr int KM = 15;   // largest fan-out (checked at creation)

#ifdef GS_DS_PROFILE
#define GS_TICK(i)                                                            \
  do {                                                                        \
    if (threadIdx.x == 0 && gridDim.x == 1) {                                 \
      const long long now_ = clock64();                                       \
      P.dbg[i] += now_ - tick_;                                               \
      tick_ = now_;                                                           \
    }                                                                         \
  } while (0)
#else
#define GS_TICK(i)
#endif

__device__ __forceinline__ uint32_t stream_raw(const uint32_t* __restrict__ R, int rem0, int next0, int64_t p) {
  return (p < rem0) ? R[next0 + p] : R[624 + (p - rem0)];
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  return y ^ (y >> 18);
}

// The container restatement of uset_emul.h with its arrays in SHARED memory (word i of thread t at smem[i*blockDim+t]:
// every lane stays in its own bank whatever it indexes).  In local memory the dynamically indexed key / next / bucket
// arrays miss L1 (most of it is carved out as shared memory here) and every one of the ~10^3 dependent accesses per
// row becomes an L2 round trip -- measured: 100 us per row batch instead of a few.
struct USetSm {
  static constexpr int kWords = 36;        // 16 keys + 4 (16 x int8 next) + 2 x 8 (32 x int8 buckets, double buffered)
  static constexpr int kNil = -1, kEmpty = -2, kBeforeBegin = -3;
  uint32_t* base;
  int stride;
  int head, nb, cnt, next_resize, boff;

  __device__ __forceinline__ USetSm(uint32_t* smem) : base(smem + threadIdx.x), stride(blockDim.x) {}
  __device__ __forceinline__ int key(int i) const { return (int)base[i * stride]; }
  __device__ __forceinline__ void set_key(int i, int v) { base[i * stride] = (uint32_t)v; }
  __device__ __forceinline__ int get8(int off, int i) const {
    const uint32_t w = base[(off + (i >> 2)) * stride];
    return (int)(int8_t)((w >> ((i & 3) * 8)) & 0xffu);
  }
  __device__ __forceinline__ void set8(int off, int i, int v) {
    uint32_t* p = base + (off + (i >> 2)) * stride;
    const int sh = (i & 3) * 8;
    *p = (*p & ~(0xffu << sh)) | (((uint32_t)v & 0xffu) << sh);
  }
  __device__ __forceinline__ int nxt(int i) const { return get8(16, i); }
  __device__ __forceinline__ void set_nxt(int i, int v) { set8(16, i, v); }
  __device__ __forceinline__ void fill_buckets(int off) {
#pragma unroll
    for (int w = 0; w < 8; ++w) base[(off + w) * stride] = 0xfefefefeu;      // kEmpty in every byte
  }
  __device__ __forceinline__ void clear() {
    nb = 1;
    cnt = 0;
    next_resize = 0;
    head = kNil;
    boff = 20;
    fill_buckets(boff);
  }
  __device__ __forceinline__ int next_of(int node) const { return node == kBeforeBegin ? head : nxt(node); }
  __device__ __forceinline__ void set_next(int node, int v) {
    if (node == kBeforeBegin) head = v;
    else set_nxt(node, v);
  }
  __device__ __forceinline__ int next_bkt(int n) {
    int r;
    if (n <= 13) r = n <= 2 ? 2 : (n == 3 ? 3 : (n <= 5 ? 5 : (n <= 7 ? 7 : (n <= 11 ? 11 : 13))));
    else r = n <= 17 ? 17 : (n <= 19 ? 19 : (n <= 23 ? 23 : (n <= 29 ? 29 : (n <= 31 ? 31 : 37))));
    next_resize = r;
    return r;
  }
  __device__ __forceinline__ void rehash(int n) {
    const int noff = boff == 20 ? 28 : 20;
    fill_buckets(noff);
    int p = head;
    head = kNil;
    int bbegin_bkt = 0;
    while (p != kNil) {
      const int nx = nxt(p);
      const int b = key(p) % n;
      const int cur = get8(noff, b);
      if (cur == kEmpty) {
        set_nxt(p, head);
        const int old_head = head;
        head = p;
        set8(noff, b, kBeforeBegin);
        if (old_head != kNil) set8(noff, bbegin_bkt, p);
        bbegin_bkt = b;
      } else {
        set_nxt(p, next_of(cur));
        set_next(cur, p);
      }
      p = nx;
    }
    boff = noff;
    nb = n;
  }
  __device__ __forceinline__ bool insert(int k) {
    bool found = false;
    for (int i = 0; i < cnt; ++i) found |= (key(i) == k);
    if (found) return false;
    if (cnt + 1 > next_resize) {
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
    set_key(node, k);
    const int cur = get8(boff, b);
    if (cur != kEmpty) {
      set_nxt(node, next_of(cur));
      set_next(cur, node);
    } else {
      set_nxt(node, head);
      if (head != kNil) set8(boff, key(head) % nb, node);
      head = node;
      set8(boff, b, kBeforeBegin);
    }
    ++cnt;
    return true;
  }
};

// phase 1 of a hop: candidates and draws per row -> offsets.  Returns (draws << 32 | candidates).
__device__ long long hop_phase1(const Geom& G, const Ptrs& P, int c, int h, int n_rows, long long* sm) {
  const int k = G.fan[h];
  const int32_t* nid = P.nid + (int64_t)c * G.lcap[G.nh];
  int32_t* rowoff = P.rowoff[h] + (int64_t)c * (G.lcap[h] + 1);
  int32_t* drawoff = P.drawoff[h] + (int64_t)c * G.lcap[h];
  const int32_t* __restrict__ rp = P.rowptr;
  const long long tot = block_scan<long long>(
      n_rows,
      [&](int t) {
        const int v = nid[t];
        const int deg = rp[v + 1] - rp[v];
        const long long cnt = deg < k ? deg : k;
        const long long dr = deg > k ? k : 0;
        return (dr << 32) | cnt;
      },
      [&](int t, long long pre, long long) {
        rowoff[t] = (int32_t)(pre & 0xffffffffll);
        drawoff[t] = (int32_t)(pre >> 32);
      },
      sm);
  if (threadIdx.x == 0) rowoff[n_rows] = (int32_t)(tot & 0xffffffffll);
  __syncthreads();
  return tot;
}

// Few rows (the hops of the serial kernel): spread them over all warps so divergent lanes serialise less.
struct RowMap {
  bool active;
  int first, step;
  __device__ __forceinline__ RowMap(int n_rows) {
    int spread = 1;
    while (spread < 32 && n_rows * spread * 2 <= (int)blockDim.x) spread *= 2;
    active = (threadIdx.x % spread) == 0;
    first = threadIdx.x / spread;
    step = blockDim.x / spread;
  }
};

// "Light" hop: what the NEXT class's stream offset needs from this hop, and nothing else.  The words a row consumes are
// fixed by the row order (phase 1); which nodes the hop discovers does not depend on the container's iteration order,
// so the chain only has to (1) turn words into Floyd positions, (2) read their column ids, (3) find the distinct new
// nodes (atomicMin marker on firstq; pos/nid stay untouched) and (4) add up (degree > k_next ? k_next : 0) over the
// rows and the new nodes = the number of words the next hop will consume.  The full hop (ordering, relabelling, block
// emission) is redone later by the class-parallel kernel, which first clears the markers through cand_u.
__device__ long long light_hop(const Geom& G, const Ptrs& P, int c, int h, int64_t draw_base, int rem0, int next0,
                               int n_rows, int k_next, long long* sm) {
  const int k = G.fan[h];
  const int32_t* nid = P.nid + (int64_t)c * G.lcap[G.nh];
  const int32_t* pos = P.pos + (int64_t)c * G.n;
  int32_t* firstq = P.firstq + (int64_t)c * G.n;
  const int32_t* rowoff = P.rowoff[h] + (int64_t)c * (G.lcap[h] + 1);
  const int32_t* drawoff = P.drawoff[h] + (int64_t)c * G.lcap[h];
  int32_t* cand_u = P.cand_u[h] + (int64_t)c * G.ncap[h];
  const int32_t* __restrict__ rp = P.rowptr;
  const int32_t* __restrict__ gcolp = P.col;
  long long mine = 0;
  const RowMap rm(n_rows);
  for (int t = rm.first; rm.active && t < n_rows; t += rm.step) {
    const int v = nid[t];
    const int q0 = rowoff[t];
    const int64_t d0 = draw_base + drawoff[t];
    const int beg = rp[v], deg = rp[v + 1] - beg;
    if (deg > k_next) mine += k_next;
    int pk[KM];
    int cnt;
    if (deg <= k) {
      cnt = deg;
#pragma unroll
      for (int i = 0; i < KM; ++i) pk[i] = i;
    } else {
      cnt = k;
      uint32_t w[KM];
#pragma unroll
      for (int i = 0; i < KM; ++i) w[i] = (i < k) ? stream_raw(P.R, rem0, next0, d0 + i) : 0u;
#pragma unroll
      for (int i = 0; i < KM; ++i) {
        pk[i] = 0;
        if (i < k) {
          const int j = deg - k + i;
          const int r = (int)(mt_temper(w[i]) % (uint32_t)j);
          bool dup = false;
#pragma unroll
          for (int e = 0; e < i; ++e) dup |= (pk[e] == r);
          pk[i] = dup ? j : r;
        }
      }
    }
    int u[KM], old[KM], ps[KM];
#pragma unroll
    for (int i = 0; i < KM; ++i) u[i] = (i < cnt) ? gcolp[beg + pk[i]] : 0;
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      old[i] = 0;
      ps[i] = 0;
      if (i < cnt) {
        old[i] = atomicMin(&firstq[u[i]], -1);
        ps[i] = pos[u[i]];
        cand_u[q0 + i] = u[i];
      }
    }
    int dg[KM];
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      const bool fresh = (i < cnt) && old[i] == INT_MAX && ps[i] < 0;
      dg[i] = fresh ? (rp[u[i] + 1] - rp[u[i]]) : 0;
    }
#pragma unroll
    for (int i = 0; i < KM; ++i)
      if (dg[i] > k_next) mine += k_next;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (threadIdx.x == 0) sm[kWarps] = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(reinterpret_cast<unsigned long long*>(&sm[kWarps]), (unsigned long long)mine);
  __syncthreads();
  const long long total = sm[kWarps];
  __syncthreads();
  return total;
}

// Full hop.  have_phase1: the row offsets (and stale light-hop markers, cleared here through cand_u) already exist.
__device__ int do_hop(const Geom& G, const Ptrs& P, int c, int h, bool record, bool have_phase1, int64_t draw_base,
                      int rem0, int next0, int n_rows, long long* draws, long long* sm, uint32_t* uset_sm) {
  const int k = G.fan[h];
  const bool last = (h == G.nh - 1);
  int32_t* nid = P.nid + (int64_t)c * G.lcap[G.nh];
  int32_t* pos = P.pos + (int64_t)c * G.n;
  int32_t* firstq = P.firstq + (int64_t)c * G.n;
  int32_t* rowoff = P.rowoff[h] + (int64_t)c * (G.lcap[h] + 1);
  int32_t* drawoff = P.drawoff[h] + (int64_t)c * G.lcap[h];
  int32_t* cand_e = P.cand_e[h] + (int64_t)c * G.ncap[h];
  int32_t* cand_u = P.cand_u[h] + (int64_t)c * G.ncap[h];
  int32_t* ocol = P.ocol[h] + (int64_t)c * G.ncap[h];
  float* oval = P.oval[h] + (int64_t)c * G.ncap[h];
  int32_t* erow = P.erow[h] + (int64_t)c * G.ncap[h];
  const int32_t* __restrict__ rp = P.rowptr;
  const int32_t* __restrict__ gcolp = P.col;
#ifdef GS_DS_PROFILE
  long long tick_ = clock64();
#endif

  int n_cand;
  if (have_phase1) {
    n_cand = rowoff[n_rows];
    *draws = 0;
    for (int q = threadIdx.x; q < n_cand; q += blockDim.x) firstq[cand_u[q]] = INT_MAX;
    __syncthreads();
  } else {
    const long long tot = hop_phase1(G, P, c, h, n_rows, sm);
    n_cand = (int)(tot & 0xffffffffll);
    *draws = tot >> 32;
  }
  GS_TICK(0);

  // phase 2: Robert-Floyd draws into the container restatement; candidates in the container's iteration order
  const RowMap rm(n_rows);
  const bool row_thread = rm.active;
  const int row_first = rm.first, row_step = rm.step;
  for (int t = row_first; row_thread && t < n_rows; t += row_step) {
    const int v = nid[t];
    const int q0 = rowoff[t];
    const int64_t d0 = draw_base + drawoff[t];
    const int beg = rp[v], deg = rp[v + 1] - beg;
    USetSm S(uset_sm);
    S.clear();
    if (deg <= k) {
      for (int j = 0; j < deg; ++j) S.insert(j);
    } else {
      uint32_t w[KM];
#pragma unroll
      for (int i = 0; i < KM; ++i) w[i] = (i < k) ? stream_raw(P.R, rem0, next0, d0 + i) : 0u;
#pragma unroll
      for (int i = 0; i < KM; ++i) {
        if (i < k) {
          const int j = deg - k + i;
          const int r = (int)(mt_temper(w[i]) % (uint32_t)j);
          if (!S.insert(r)) S.insert(j);
        }
      }
    }
    int keys[KM];
    const int cnt = S.cnt;
    int p = S.head;
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      keys[i] = 0;
      if (i < cnt) {
        keys[i] = S.key(p);
        p = S.nxt(p);
      }
    }
    int u[KM], ps[KM];
#pragma unroll
    for (int i = 0; i < KM; ++i) u[i] = (i < cnt) ? gcolp[beg + keys[i]] : 0;
#pragma unroll
    for (int i = 0; i < KM; ++i) ps[i] = (i < cnt) ? pos[u[i]] : 0;
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      if (i < cnt) {
        cand_e[q0 + i] = beg + keys[i];
        cand_u[q0 + i] = u[i];
        if (ps[i] < 0) atomicMin(&firstq[u[i]], q0 + i);
      }
    }
  }
  __syncthreads();
  GS_TICK(1);

  // phase 3: first occurrences get the next class-local ids, in candidate order
  const long long n_new = block_scan<long long>(
      n_cand,
      [&](int q) {
        const int u = cand_u[q];
        const int pu = pos[u], fq = firstq[u];
        return (long long)((pu < 0 && fq == q) ? 1 : 0);
      },
      [&](int q, long long pre, long long isnew) {
        if (isnew) {
          const int u = cand_u[q];
          const int loc = n_rows + (int)pre;
          nid[loc] = u;
          pos[u] = loc;
        }
      },
      sm);
  __syncthreads();
  GS_TICK(2);

  // phase 4: relabel and emit the class block with every row sorted by local column.  The locals of a row are
  // distinct, so an entry's place is its rank among the row's locals (all-pairs compare, no data-dependent loop).
  for (int t = row_first; row_thread && t < n_rows; t += row_step) {
    const int q0 = rowoff[t], cnt = rowoff[t + 1] - q0;
    int cu[KM], ce[KM], loc[KM];
    float vv[KM];
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      cu[i] = (i < cnt) ? cand_u[q0 + i] : 0;
      ce[i] = (i < cnt) ? cand_e[q0 + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      loc[i] = (i < cnt) ? pos[cu[i]] : INT_MAX;
      vv[i] = (i < cnt && record) ? P.val[ce[i]] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < KM; ++i)
      if (i < cnt) firstq[cu[i]] = INT_MAX;
    if (!record) continue;
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      if (i < cnt) {
        int rank = 0;
#pragma unroll
        for (int j = 0; j < KM; ++j) rank += (loc[j] < loc[i]) ? 1 : 0;
        ocol[q0 + rank] = loc[i];
        oval[q0 + rank] = vv[i];
        erow[q0 + rank] = t;
        if (last) ce[i] = rank;                    // remember where the entry went (cand_u is rewritten below)
      }
    }
    if (last) {                                    // global column ids of the outermost hop (fused feature gather)
#pragma unroll
      for (int i = 0; i < KM; ++i)
        if (i < cnt) cand_u[q0 + ce[i]] = cu[i];
    }
  }
  __syncthreads();
  GS_TICK(3);
  return n_rows + (int)n_new;
}

// ------------------------------------------------------------------------------------------ per-class preparation
// Class-parallel and independent of the random stream: batch -> n_id / pos map, and (two-hop case) the row offsets of
// hop 0, which only depend on degrees.  For a single hop the only thing the serial kernel needs is the draw count.
__global__ void __launch_bounds__(kThreads, 1)
sample_prep_kernel(Geom G, Ptrs P, int n_class, const int32_t* __restrict__ batch,
                   const int32_t* __restrict__ batch_off) {
  __shared__ long long sm[kWarps + 2];
  const int c = blockIdx.x;
  if (c >= n_class) return;
  int32_t* nid = P.nid + (int64_t)c * G.lcap[G.nh];
  int32_t* pos = P.pos + (int64_t)c * G.n;
  int32_t* lc = P.level_count + (int64_t)c * (G.nh + 1);
  const int b0 = batch_off[c];
  const int n_rows = batch_off[c + 1] - b0;
  const int k_last = G.fan[G.nh - 1];
  long long mine = 0;
  for (int t = threadIdx.x; t < n_rows; t += blockDim.x) {
    const int v = batch[b0 + t];
    nid[t] = v;
    pos[v] = t;
    if (G.nh == 1 && P.rowptr[v + 1] - P.rowptr[v] > k_last) mine += k_last;
  }
  if (threadIdx.x == 0) lc[0] = n_rows;
  __syncthreads();
  if (G.nh == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (threadIdx.x == 0) sm[kWarps] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(reinterpret_cast<unsigned long long*>(&sm[kWarps]), (unsigned long long)mine);
    __syncthreads();
    if (threadIdx.x == 0) P.hop_tot[c] = sm[kWarps];
  } else if (G.nh == 2) {
    const long long tot = hop_phase1(G, P, c, 0, n_rows, sm);
    if (threadIdx.x == 0) P.hop_tot[c] = tot;
  }
}

// ------------------------------------------------------------------------------------------ serial part (1 CTA)
// Cuts the random stream into per-class, per-hop segments.  Per class: full hops 0..nh-3 (only for three or more
// hops: their exact n_id order fixes the next hop's row offsets), then the light version of hop nh-2, whose result is
// the number of words hop nh-1 will consume.  Hops nh-2 and nh-1 themselves run class-parallel afterwards.
__global__ void __launch_bounds__(kThreads, 1)
sample_serial_kernel(Geom G, Ptrs P, int n_class, const uint8_t* __restrict__ materialise) {
  extern __shared__ uint32_t uset_sm[];
  __shared__ long long sm[kWarps + 2];
#ifdef GS_DS_PROFILE
  const long long t_begin = clock64();
#endif
  const int rem0 = (int)P.step_info[1], next0 = (int)P.step_info[2];
  const int nh = G.nh;
  long long draw_base = 0;
  for (int c = 0; c < n_class; ++c) {
    const bool keep = materialise == nullptr || materialise[c] != 0;
    int32_t* lc = P.level_count + (int64_t)c * (nh + 1);
    int n_rows = lc[0];
    if (nh == 1) {
      if (threadIdx.x == 0) P.last_off[c] = draw_base;
      draw_base += P.hop_tot[c];
      continue;
    }
    for (int h = 0; h + 2 < nh; ++h) {
      long long dr;
      n_rows = do_hop(G, P, c, h, keep, false, draw_base, rem0, next0, n_rows, &dr, sm, uset_sm);
      draw_base += dr;
      if (threadIdx.x == 0) lc[h + 1] = n_rows;
    }
    const long long tot = (nh == 2) ? P.hop_tot[c] : hop_phase1(G, P, c, nh - 2, n_rows, sm);
    if (threadIdx.x == 0) P.hop_off[c] = draw_base;
    const long long next_draws = light_hop(G, P, c, nh - 2, draw_base, rem0, next0, n_rows, G.fan[nh - 1], sm);
    draw_base += tot >> 32;
    if (threadIdx.x == 0) P.last_off[c] = draw_base;
    draw_base += next_draws;
  }
  // generator state after `draw_base` draws (torch's mt19937: `left` counts down from 624, reload when it hits 0)
  const long long T = draw_base;
  long long blk, left, next;
  if (T <= rem0) {
    blk = 0;
    next = next0 + T;
    left = (rem0 + 1) - T;
  } else {
    const long long q = T - rem0;            // reads beyond block 0, q >= 1
    blk = 1 + (q - 1) / 624;
    const long long i = (q - 1) % 624;       // index of the last word read
    next = i + 1;
    left = 624 - i;
  }
  for (int i = threadIdx.x; i < 624; i += blockDim.x) P.mt->s[i] = P.R[blk * 624 + i];
  if (threadIdx.x == 0) {
    P.mt->left = (int32_t)left;
    P.mt->next = (int32_t)next;
    P.step_info[0] = T;
#ifdef GS_DS_PROFILE
    P.dbg[8] += clock64() - t_begin;
#endif
  }
}

// ------------------------------------------------------------------------------------------ last hop, 1 CTA / class
// transposed class block: rows = class-local columns, entries in source-row order (what a stable sort by column gives).
// Placement is inherently ordered, so ONE warp places 32 entries per step (__match_any_sync ranks equal columns, the
// column cursors live in shared memory); the other warps stage the next 1024 entries into shared memory meanwhile, so a
// step costs shared-memory latency only.
constexpr int kTile = 1024;
__device__ void build_transpose(const Geom& G, const Ptrs& P, int c, int h, int n_rows, int n_cols, int32_t* cursor_sm,
                                int32_t* tile_sm, long long* sm) {
  const int32_t* rowoff = P.rowoff[h] + (int64_t)c * (G.lcap[h] + 1);
  const int32_t* ocol = P.ocol[h] + (int64_t)c * G.ncap[h];
  const float* oval = P.oval[h] + (int64_t)c * G.ncap[h];
  const int32_t* erow = P.erow[h] + (int64_t)c * G.ncap[h];
  int32_t* tr = P.t_rowptr[h] + (int64_t)c * (G.lcap[h + 1] + 1);
  int32_t* tc = P.t_col[h] + (int64_t)c * G.ncap[h];
  float* tv = P.t_val[h] + (int64_t)c * G.ncap[h];
  int32_t* cursor = (n_cols <= kCursorSmemInts) ? cursor_sm : P.t_cursor + (int64_t)c * G.lcap[G.nh];
  const int nnz = rowoff[n_rows];
  for (int j = threadIdx.x; j <= n_cols; j += blockDim.x) tr[j] = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < nnz; e += blockDim.x) atomicAdd(&tr[ocol[e] + 1], 1);
  __syncthreads();
  block_scan<long long>(
      n_cols, [&](int j) { return (long long)tr[j + 1]; },
      [&](int j, long long pre, long long) { cursor[j] = (int32_t)pre; }, sm);
  __syncthreads();
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) tr[j] = cursor[j];
  if (threadIdx.x == 0) tr[n_cols] = nnz;
  __syncthreads();
  const int ntiles = (nnz + kTile - 1) / kTile;
  auto stage = [&](int tl, int first_thread) {
    int32_t* bc = tile_sm + (tl & 1) * 3 * kTile;
    int32_t* br = bc + kTile;
    float* bv = reinterpret_cast<float*>(br + kTile);
    for (int i = threadIdx.x - first_thread; i < kTile; i += blockDim.x - first_thread) {
      const int e = tl * kTile + i;
      if (e < nnz) {
        bc[i] = ocol[e];
        br[i] = erow[e];
        bv[i] = oval[e];
      }
    }
  };
  if (ntiles > 0) stage(0, 0);
  __syncthreads();
  for (int tl = 0; tl < ntiles; ++tl) {
    if (threadIdx.x >= 32) {
      if (tl + 1 < ntiles) stage(tl + 1, 32);
    } else {
      const int lane = threadIdx.x;
      const int32_t* bc = tile_sm + (tl & 1) * 3 * kTile;
      const int32_t* br = bc + kTile;
      const float* bv = reinterpret_cast<const float*>(br + kTile);
      const int m_tile = min(kTile, nnz - tl * kTile);
      for (int i0 = 0; i0 < m_tile; i0 += 32) {
        const int i = i0 + lane;
        const bool live = i < m_tile;
        const int cj = live ? bc[i] : -1 - lane;
        const unsigned m = __match_any_sync(0xffffffffu, cj);
        const int rank = __popc(m & ((1u << lane) - 1u));
        int base = 0;
        if (live) base = cursor[cj];
        __syncwarp();
        if (live && rank == 0) cursor[cj] = base + __popc(m);
        __syncwarp();
        if (live) {
          tc[base + rank] = br[i];
          tv[base + rank] = bv[i];
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads, 1)
sample_last_kernel(Geom G, Ptrs P, int n_class, const uint8_t* __restrict__ materialise) {
  extern __shared__ int32_t dyn_sm[];
  int32_t* cursor_sm = dyn_sm;
  int32_t* tile_sm = dyn_sm + kCursorSmemInts;
  __shared__ long long sm[kWarps + 2];
  const int c = blockIdx.x;
  if (c >= n_class) return;
  const bool keep = materialise == nullptr || materialise[c] != 0;
  const int rem0 = (int)P.step_info[1], next0 = (int)P.step_info[2];
  int32_t* nid = P.nid + (int64_t)c * G.lcap[G.nh];
  int32_t* pos = P.pos + (int64_t)c * G.n;
  int32_t* lc = P.level_count + (int64_t)c * (G.nh + 1);
  const int nh = G.nh;
  uint32_t* uset_sm = reinterpret_cast<uint32_t*>(dyn_sm);      // the cursor region is idle during the hops
  int n_rows = lc[nh >= 2 ? nh - 2 : 0];
  const int n_pos = n_rows;                                     // nodes whose pos entry is set on entry
  if (keep) {
    long long dr;
    if (nh >= 2) {
      n_rows = do_hop(G, P, c, nh - 2, true, true, P.hop_off[c], rem0, next0, n_rows, &dr, sm, uset_sm);
      if (threadIdx.x == 0) lc[nh - 1] = n_rows;
      __syncthreads();
    }
    n_rows = do_hop(G, P, c, nh - 1, true, false, P.last_off[c], rem0, next0, n_rows, &dr, sm, uset_sm);
    if (threadIdx.x == 0) lc[nh] = n_rows;
    __syncthreads();
    for (int h = 0; h < nh; ++h) build_transpose(G, P, c, h, lc[h], lc[h + 1], cursor_sm, tile_sm, sm);
    for (int t = threadIdx.x; t < n_rows; t += blockDim.x) pos[nid[t]] = -1;
  } else {
    if (nh >= 2) {                                              // clear the light hop's markers
      const int32_t* rowoff = P.rowoff[nh - 2] + (int64_t)c * (G.lcap[nh - 2] + 1);
      const int32_t* cand_u = P.cand_u[nh - 2] + (int64_t)c * G.ncap[nh - 2];
      int32_t* firstq = P.firstq + (int64_t)c * G.n;
      const int n_cand = rowoff[n_rows];
      for (int q = threadIdx.x; q < n_cand; q += blockDim.x) firstq[cand_u[q]] = INT_MAX;
    }
    for (int t = threadIdx.x; t < n_pos; t += blockDim.x) pos[nid[t]] = -1;
  }
}

// ------------------------------------------------------------------------------------------ packing
// Same layout and `desc` table as gs_sampler_finish_step (host_sampler.cpp); offsets are byte offsets into `out`.
__global__ void __launch_bounds__(kThreads)
pack_offsets_kernel(Geom G, Ptrs P, int n_class, const uint8_t* __restrict__ materialise, int has_labels,
                    int64_t out_cap, int64_t* __restrict__ desc) {
  __shared__ long long sm[kWarps + 2];
  const int nh = G.nh, al = G.align;
  for (int l = 0; l <= nh; ++l) {
    int32_t* s = P.seg + (int64_t)l * (n_class + 1);
    const long long tot = block_scan<long long>(
        n_class,
        [&](int c) {
          const bool keep = materialise == nullptr || materialise[c] != 0;
          const int cnt = keep ? P.level_count[(int64_t)c * (nh + 1) + l] : 0;
          return (long long)((cnt + al - 1) / al * al);
        },
        [&](int c, long long pre, long long) { s[c] = (int32_t)pre; }, sm);
    if (threadIdx.x == 0) s[n_class] = (int32_t)tot;
  }
  for (int h = 0; h < nh; ++h) {
    int64_t* e = P.eoff + (int64_t)h * (n_class + 1);
    const long long tot = block_scan<long long>(
        n_class,
        [&](int c) {
          const bool keep = materialise == nullptr || materialise[c] != 0;
          const int rows = P.level_count[(int64_t)c * (nh + 1) + h];
          return keep ? (long long)(P.rowoff[h] + (int64_t)c * (G.lcap[h] + 1))[rows] : 0ll;
        },
        [&](int c, long long pre, long long) { e[c] = pre; }, sm);
    if (threadIdx.x == 0) e[n_class] = tot;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int i = 0; i < 64; ++i) desc[i] = -1;
  int64_t off = 0;
  bool ok = true;
  auto reserve = [&](int64_t bytes) -> int64_t {
    const int64_t at = off;
    if (at + bytes > out_cap) ok = false;
    off = (at + bytes + 15) & ~int64_t(15);
    return at;
  };
  desc[0] = nh;
  desc[1] = P.seg[n_class];
  for (int l = 0; l <= nh; ++l) desc[2 + l] = P.seg[(int64_t)l * (n_class + 1) + n_class];
  const int64_t n_tgt = desc[2], n_last = desc[2 + nh];
  desc[8] = reserve((int64_t)(nh + 1) * (n_class + 1) * 4);
  desc[9] = reserve(n_last * 4);
  desc[10] = reserve(n_tgt * 4);
  desc[11] = reserve(n_tgt * 4);
  desc[12] = reserve(n_tgt * 4);
  if (has_labels) desc[13] = reserve(n_tgt * 4);
  desc[14] = reserve((int64_t)(nh + 1) * n_class * 4);
  for (int h = 0; h < nh; ++h) {
    const int64_t n_rows = desc[2 + h], n_cols = desc[3 + h];
    const int64_t nnz = P.eoff[(int64_t)h * (n_class + 1) + n_class];
    int64_t* d = desc + 16 + 8 * h;
    d[0] = nnz;
    d[1] = reserve((n_rows + 1) * 4);
    d[2] = reserve(nnz * 4);
    d[3] = reserve(nnz * 4);
    d[4] = reserve((n_cols + 1) * 4);
    d[5] = reserve(nnz * 4);
    d[6] = reserve(nnz * 4);
    if (h == nh - 1) d[7] = reserve(nnz * 4);
  }
  desc[60] = P.step_info[0];     // draws consumed
  desc[62] = off;                // bytes used
  desc[63] = ok ? 0 : GS_ENOSPC;
}

constexpr int kPackSplit = 8;    // CTAs per class in pack_copy_kernel (blockIdx.y)
__global__ void __launch_bounds__(256)
pack_copy_kernel(Geom G, Ptrs P, int n_class, const uint8_t* __restrict__ materialise, uint8_t* __restrict__ out,
                 const int64_t* __restrict__ desc) {
  const int c = blockIdx.x;
  const int nh = G.nh;
  if (desc[63] != 0) return;
  const bool keep = materialise == nullptr || materialise[c] != 0;
  const int tid = threadIdx.x + blockDim.x * blockIdx.y, nt = blockDim.x * gridDim.y;
  const int32_t* lc = P.level_count + (int64_t)c * (nh + 1);
  int32_t* cnt = reinterpret_cast<int32_t*>(out + desc[14]);
  for (int l = tid; l <= nh; l += nt) cnt[(int64_t)l * n_class + c] = keep ? lc[l] : 0;
  if (c == 0) {
    int32_t* segs = reinterpret_cast<int32_t*>(out + desc[8]);
    for (int i = tid; i < (nh + 1) * (n_class + 1); i += nt) segs[i] = P.seg[i];
    for (int h = tid; h < nh; h += nt) {
      reinterpret_cast<int32_t*>(out + desc[16 + 8 * h + 1])[0] = 0;
      reinterpret_cast<int32_t*>(out + desc[16 + 8 * h + 4])[0] = 0;
    }
  }
  if (!keep) return;
  const int32_t* nid = P.nid + (int64_t)c * G.lcap[nh];
  auto seg = [&](int l, int cc) { return P.seg[(int64_t)l * (n_class + 1) + cc]; };
  {
    int32_t* nid_all = reinterpret_cast<int32_t*>(out + desc[9]) + seg(nh, c);
    const int m = lc[nh], mp = seg(nh, c + 1) - seg(nh, c);
    for (int t = tid; t < mp; t += nt) nid_all[t] = t < m ? nid[t] : 0;
    const int o = seg(0, c), bsz = lc[0], bp = seg(0, c + 1) - o;
    int32_t* tcls = reinterpret_cast<int32_t*>(out + desc[10]) + o;
    float* inv_b = reinterpret_cast<float*>(out + desc[11]) + o;
    int32_t* tgt = reinterpret_cast<int32_t*>(out + desc[12]) + o;
    int32_t* tlab = desc[13] >= 0 ? reinterpret_cast<int32_t*>(out + desc[13]) + o : nullptr;
    const float w = 1.0f / (float)bsz;
    for (int t = tid; t < bp; t += nt) {
      const bool real = t < bsz;
      tcls[t] = c;
      inv_b[t] = real ? w : 0.0f;
      tgt[t] = real ? nid[t] : 0;
      if (tlab) tlab[t] = real ? P.labels[nid[t]] : 0;
    }
  }
  for (int h = 0; h < nh; ++h) {
    const int64_t* d = desc + 16 + 8 * h;
    const int64_t e0 = P.eoff[(int64_t)h * (n_class + 1) + c];
    const int rows_c = lc[h], cols_c = lc[h + 1];
    const int r0 = seg(h, c), r1 = seg(h, c + 1), c0 = seg(h + 1, c), c1 = seg(h + 1, c + 1);
    const int32_t* rowoff = P.rowoff[h] + (int64_t)c * (G.lcap[h] + 1);
    const int m = rowoff[rows_c];
    int32_t* rowptr = reinterpret_cast<int32_t*>(out + d[1]);
    for (int r = tid; r < r1 - r0; r += nt) rowptr[r0 + r + 1] = (int32_t)(e0 + (r < rows_c ? rowoff[r + 1] : m));
    int32_t* col = reinterpret_cast<int32_t*>(out + d[2]) + e0;
    float* val = reinterpret_cast<float*>(out + d[3]) + e0;
    const int32_t* ocol = P.ocol[h] + (int64_t)c * G.ncap[h];
    const float* oval = P.oval[h] + (int64_t)c * G.ncap[h];
    for (int i = tid; i < m; i += nt) {
      col[i] = ocol[i] + c0;
      val[i] = oval[i];
    }
    if (h == nh - 1) {
      int32_t* gcol = reinterpret_cast<int32_t*>(out + d[7]) + e0;
      const int32_t* gsrc = P.cand_u[h] + (int64_t)c * G.ncap[h];
      for (int i = tid; i < m; i += nt) gcol[i] = gsrc[i];
    }
    const int32_t* tr = P.t_rowptr[h] + (int64_t)c * (G.lcap[h + 1] + 1);
    int32_t* t_rowptr = reinterpret_cast<int32_t*>(out + d[4]);
    for (int j = tid; j < c1 - c0; j += nt) t_rowptr[c0 + j + 1] = (int32_t)(e0 + (j < cols_c ? tr[j + 1] : m));
    int32_t* t_col = reinterpret_cast<int32_t*>(out + d[5]) + e0;
    float* t_val = reinterpret_cast<float*>(out + d[6]) + e0;
    const int32_t* tcs = P.t_col[h] + (int64_t)c * G.ncap[h];
    const float* tvs = P.t_val[h] + (int64_t)c * G.ncap[h];
    for (int i = tid; i < m; i += nt) {
      t_col[i] = tcs[i] + r0;
      t_val[i] = tvs[i];
    }
  }
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace ds
}  // namespace gs

// ---------------------------------------------------------------------------------------------- C ABI
struct gs_dsampler {
  gs::ds::Geom G;
  gs::ds::Ptrs P;
  void* arena = nullptr;
  int64_t arena_bytes = 0;
  int64_t r_blocks = 0;
  int64_t out_cap = 0;
  bool has_labels = false;
};

extern "C" {
using namespace gs;
using namespace gs::ds;

static int64_t dsampler_max_draws(const Geom& G, int64_t n_class, int64_t batch) {
  int64_t rows = batch, draws = 0;
  for (int h = 0; h < G.nh; ++h) {
    draws += rows * G.fan[h];
    rows += rows * G.fan[h];
  }
  return draws * n_class;
}

gs_dsampler* gs_dsampler_create(int32_t n_nodes, const int32_t* d_rowptr, const int32_t* d_col, const float* d_val,
                                const int32_t* d_labels, int32_t n_hops, const int32_t* fanout, int32_t n_class_max,
                                int32_t batch_max, int32_t align, void* stream) {
  if (n_nodes < 1 || !d_rowptr || !d_col || !d_val || n_hops < 1 || n_hops > kMaxHops || !fanout || n_class_max < 1 ||
      batch_max < 1 || align < 1) {
    set_error_msg("gs_dsampler_create: invalid argument");
    return nullptr;
  }
  for (int i = 0; i < n_hops; ++i)
    if (fanout[i] < 0 || fanout[i] > 15) {
      set_error_msg("gs_dsampler_create: fan-outs above 15 are not supported (container restatement is sized for 16 keys)");
      return nullptr;
    }
  gs_dsampler* S = new gs_dsampler();
  Geom& G = S->G;
  G.n = n_nodes;
  G.nh = n_hops;
  G.n_class_max = n_class_max;
  G.batch_max = batch_max;
  G.align = align;
  G.lcap[0] = batch_max;
  for (int h = 0; h < n_hops; ++h) {
    G.fan[h] = fanout[h];
    G.ncap[h] = G.lcap[h] * fanout[h];
    // a level cannot hold more nodes than the graph has
    G.lcap[h + 1] = std::min<int64_t>(G.lcap[h] + G.ncap[h], (int64_t)n_nodes);
  }
  const int64_t nc = n_class_max;
  S->r_blocks = (dsampler_max_draws(G, nc, batch_max) + 623) / 624 + 2;
  // one arena, 256-byte aligned pieces
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    const int64_t at = off;
    off = (off + bytes + 255) & ~int64_t(255);
    return at;
  };
  struct Piece {
    void** dst;
    int64_t at;
  };
  std::vector<Piece> pieces;
  auto want = [&](void** dst, int64_t bytes) { pieces.push_back(Piece{dst, take(bytes)}); };
  Ptrs& P = S->P;
  P.rowptr = d_rowptr;
  P.col = d_col;
  P.val = d_val;
  P.labels = d_labels;
  S->has_labels = d_labels != nullptr;
  want((void**)&P.mt, sizeof(MtDev));
  want((void**)&P.R, S->r_blocks * 624 * 4);
  want((void**)&P.step_info, 8 * 8);
  want((void**)&P.pos, nc * (int64_t)n_nodes * 4);
  want((void**)&P.firstq, nc * (int64_t)n_nodes * 4);
  want((void**)&P.nid, nc * G.lcap[n_hops] * 4);
  want((void**)&P.level_count, nc * (n_hops + 1) * 4);
  want((void**)&P.last_off, nc * 8);
  want((void**)&P.hop_off, nc * 8);
  want((void**)&P.hop_tot, nc * 8);
  for (int h = 0; h < n_hops; ++h) {
    want((void**)&P.rowoff[h], nc * (G.lcap[h] + 1) * 4);
    want((void**)&P.drawoff[h], nc * G.lcap[h] * 4);
    want((void**)&P.cand_e[h], nc * G.ncap[h] * 4);
    want((void**)&P.cand_u[h], nc * G.ncap[h] * 4);
    want((void**)&P.ocol[h], nc * G.ncap[h] * 4);
    want((void**)&P.oval[h], nc * G.ncap[h] * 4);
    want((void**)&P.erow[h], nc * G.ncap[h] * 4);
    want((void**)&P.t_rowptr[h], nc * (G.lcap[h + 1] + 1) * 4);
    want((void**)&P.t_col[h], nc * G.ncap[h] * 4);
    want((void**)&P.t_val[h], nc * G.ncap[h] * 4);
  }
  want((void**)&P.t_cursor, nc * G.lcap[n_hops] * 4);
  want((void**)&P.seg, (int64_t)(n_hops + 1) * (nc + 1) * 4);
  want((void**)&P.eoff, (int64_t)n_hops * (nc + 1) * 8);
  want((void**)&P.dbg, 16 * 8);
  S->arena_bytes = off;
  cudaError_t e = cudaMalloc(&S->arena, (size_t)off);
  if (e != cudaSuccess) {
    set_error("gs_dsampler_create: cudaMalloc", e);
    delete S;
    return nullptr;
  }
  for (const Piece& p : pieces) *p.dst = static_cast<uint8_t*>(S->arena) + p.at;
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(P.dbg, 0, 16 * 8, st);
  fill_i32_kernel<<<kNumSMs * 4, 256, 0, st>>>(P.pos, nc * (int64_t)n_nodes, -1);
  finish_launch("dsampler_fill_pos");
  fill_i32_kernel<<<kNumSMs * 4, 256, 0, st>>>(P.firstq, nc * (int64_t)n_nodes, INT_MAX);
  finish_launch("dsampler_fill_firstq");
  // packed output capacity: every class segment padded to `align` rows at every level
  {
    auto pad = [&](int64_t x) { return (x + align - 1) / align * align; };
    int64_t cap = 4 * (int64_t)(n_hops + 1) * (nc + 1) + 64;
    int64_t lvl = pad(batch_max) * nc;
    const int64_t rows0 = lvl;
    for (int h = 0; h < n_hops; ++h) {
      const int64_t nnz = lvl * fanout[h];
      cap += 4 * (lvl + 1) + 7 * 4 * nnz + 4 * (lvl + nnz + (int64_t)align * nc + 1) + 8 * 16;
      lvl = lvl + nnz + (int64_t)align * nc;
    }
    cap += 4 * lvl + 4 * 4 * rows0 + 4 * (int64_t)(n_hops + 1) * nc + 16 * 8;
    S->out_cap = cap;
  }
  if (cudaFuncSetAttribute(sample_serial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           USetSm::kWords * kThreads * 4) != cudaSuccess) {
    cudaGetLastError();
    set_error_msg("gs_dsampler_create: cannot reserve shared memory for the serial sampling kernel");
    cudaFree(S->arena);
    delete S;
    return nullptr;
  }
  if (cudaFuncSetAttribute(sample_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (kCursorSmemInts + 2 * 3 * kTile) * 4) !=
      cudaSuccess) {
    cudaGetLastError();
    set_error_msg("gs_dsampler_create: cannot reserve shared memory for the transposed-block cursors");
    cudaFree(S->arena);
    delete S;
    return nullptr;
  }
  return S;
}

void gs_dsampler_destroy(gs_dsampler* S) {
  if (!S) return;
  if (S->arena) cudaFree(S->arena);
  delete S;
}

int64_t gs_dsampler_out_capacity(const gs_dsampler* S) { return S ? S->out_cap : GS_EINVAL; }
int64_t gs_dsampler_scratch_bytes(const gs_dsampler* S) { return S ? S->arena_bytes : GS_EINVAL; }

// torch's mt19937 engine state (624 words, `left`, `next`) -> device
int gs_dsampler_set_rng(gs_dsampler* S, const uint32_t* state, int32_t left, int32_t next, void* stream) {
  GS_REQUIRE(S && state && left >= 1 && left <= 624 && next >= 0 && next <= 624);
  MtDev h;
  std::memcpy(h.s, state, sizeof(h.s));
  h.left = left;
  h.next = next;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemcpyAsync(S->P.mt, &h, sizeof(h), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);     // `h` is a stack object
  if (e != cudaSuccess) {
    set_error("gs_dsampler_set_rng", e);
    return (int)e;
  }
  return GS_OK;
}

// device -> host after all sampling work queued on `stream` (synchronises the stream)
int gs_dsampler_get_rng(gs_dsampler* S, uint32_t* state, int32_t* left, int32_t* next, void* stream) {
  GS_REQUIRE(S && state && left && next);
  MtDev h;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemcpyAsync(&h, S->P.mt, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("gs_dsampler_get_rng", e);
    return (int)e;
  }
  std::memcpy(state, h.s, sizeof(h.s));
  *left = h.left;
  *next = h.next;
  return GS_OK;
}

// One outer step.  d_batch: concatenated class batches (device, int32 node ids); d_batch_off: n_class + 1 offsets
// (device); d_materialise: per class 0/1 (device) or NULL = all; max_batch: largest class batch (host value, bounds the
// random words generated ahead).  Everything is queued on `stream`; d_desc (64 x int64, device) receives the layout of
// the packed blocks in d_out exactly as gs_sampler_finish_step reports it, plus desc[60] = draws, desc[62] = bytes used,
// desc[63] = 0 or GS_ENOSPC.
int gs_dsampler_sample_step(gs_dsampler* S, int32_t n_class, const int32_t* d_batch, const int32_t* d_batch_off,
                            const uint8_t* d_materialise, int32_t max_batch, uint8_t* d_out, int64_t out_cap,
                            int64_t* d_desc, void* stream) {
  GS_REQUIRE(S && d_batch && d_batch_off && d_out && d_desc && n_class >= 1 && n_class <= S->G.n_class_max &&
             max_batch >= 0 && max_batch <= S->G.batch_max && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  const Geom& G = S->G;
  int64_t nblocks = (dsampler_max_draws(G, n_class, max_batch) + 623) / 624 + 1;
  if (nblocks > S->r_blocks - 1) nblocks = S->r_blocks - 1;
  mt_generate_kernel<<<1, 256, 0, st>>>(S->P.mt, S->P.R, (int)nblocks, S->P.step_info);
  int rc = finish_launch("dsampler_mt_generate");
  if (rc) return rc;
  sample_prep_kernel<<<n_class, kThreads, 0, st>>>(G, S->P, n_class, d_batch, d_batch_off);
  if ((rc = finish_launch("dsampler_prep"))) return rc;
  sample_serial_kernel<<<1, kThreads, USetSm::kWords * kThreads * 4, st>>>(G, S->P, n_class, d_materialise);
  if ((rc = finish_launch("dsampler_serial"))) return rc;
  sample_last_kernel<<<n_class, kThreads, (kCursorSmemInts + 2 * 3 * kTile) * 4, st>>>(G, S->P, n_class, d_materialise);
  if ((rc = finish_launch("dsampler_last_hop"))) return rc;
  pack_offsets_kernel<<<1, kThreads, 0, st>>>(G, S->P, n_class, d_materialise, S->has_labels ? 1 : 0, out_cap, d_desc);
  if ((rc = finish_launch("dsampler_pack_offsets"))) return rc;
  pack_copy_kernel<<<dim3(n_class, kPackSplit), 256, 0, st>>>(G, S->P, n_class, d_materialise, d_out, d_desc);
  return finish_launch("dsampler_pack_copy");
}

// Host-side check of the container restatement (tests): iteration order after inserting keys[0..n).
int64_t gs_uset_emul_order(const int64_t* keys, int64_t n, int64_t* out) {
  if (!keys || !out || n < 0) return GS_EINVAL;
  gs::USetEmul s;
  s.clear();
  for (int64_t i = 0; i < n; ++i) {
    if (s.cnt >= gs::USetEmul::kMax) return GS_ENOSPC;
    s.insert((int32_t)keys[i]);
  }
  int64_t m = 0;
  for (int p = s.head; p != gs::USetEmul::kNil; p = s.nxt[p]) out[m++] = s.key[p];
  return m;
}

}  // extern "C"

extern "C" int gs_dsampler_debug_counters(gs_dsampler* S, int64_t* out16, void* stream) {
  GS_REQUIRE(S && out16);
  cudaStream_t st = gs::as_stream(stream);
  cudaError_t e = cudaMemcpyAsync(out16, S->P.dbg, 16 * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e == cudaSuccess ? GS_OK : (int)e;
}
