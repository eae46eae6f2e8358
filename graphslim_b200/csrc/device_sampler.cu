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
//     phases each), so draw number p of the step is a plain array read (stream_word);
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
constexpr int kThreads = 1024;
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
};

// ------------------------------------------------------------------------------------------ mt19937 stream
__device__ __forceinline__ uint32_t mt_tw(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

// R[0] <- current state; R[b] <- regeneration of R[b-1].  One CTA; the three phases of a block only depend on the
// previous phase (new[i] = new[i-227] ^ tw(old[i], old[i+1]) for i >= 227), so a block costs four barriers.
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
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t v = nw[396] ^ mt_tw(o[623], nw[0]);
      nw[623] = v;
      dst[623] = v;
    }
    __syncthreads();
  }
}

// tempered output number p (0-based) of the step's stream
__device__ __forceinline__ uint32_t stream_word(const uint32_t* __restrict__ R, int rem0, int next0, int64_t p) {
  uint32_t y = (p < rem0) ? R[next0 + p] : R[624 + (p - rem0)];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  return y ^ (y >> 18);
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
      sm[lane] = w;
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
__device__ int do_hop(const Geom& G, const Ptrs& P, int c, int h, bool record, int64_t draw_base, int rem0, int next0,
                      int n_rows, long long* draws, long long* sm) {
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

  // phase 1: candidates and draws per row -> offsets (high word: draws, low word: candidates)
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
  const int n_cand = (int)(tot & 0xffffffffll);
  *draws = tot >> 32;
  if (threadIdx.x == 0) rowoff[n_rows] = n_cand;
  __syncthreads();

  // phase 2: Robert-Floyd draws into the container restatement; candidates in the container's iteration order
  for (int t = threadIdx.x; t < n_rows; t += blockDim.x) {
    const int v = nid[t];
    const int beg = rp[v], deg = rp[v + 1] - beg;
    USetEmul S;
    S.clear();
    if (deg <= k) {
      for (int j = 0; j < deg; ++j) S.insert(j);
    } else {
      const int64_t d0 = draw_base + drawoff[t];
      for (int j = deg - k; j < deg; ++j) {
        const uint32_t w = stream_word(P.R, rem0, next0, d0 + (j - (deg - k)));
        const int r = (int)(w % (uint32_t)j);
        if (!S.insert(r)) S.insert(j);
      }
    }
    int q = rowoff[t];
    for (int p = S.head; p != USetEmul::kNil; p = S.nxt[p]) {
      const int e = beg + S.key[p];
      const int u = gcolp[e];
      cand_e[q] = e;
      cand_u[q] = u;
      if (pos[u] < 0) atomicMin(&firstq[u], q);
      ++q;
    }
  }
  __syncthreads();

  // phase 3: first occurrences get the next class-local ids, in candidate order
  const long long n_new = block_scan<long long>(
      n_cand,
      [&](int q) {
        const int u = cand_u[q];
        return (long long)((pos[u] < 0 && firstq[u] == q) ? 1 : 0);
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

  // phase 4: relabel, sort every row by local column, emit the class block
  for (int t = threadIdx.x; t < n_rows; t += blockDim.x) {
    const int q0 = rowoff[t], cnt = rowoff[t + 1] - q0;
    int loc[USetEmul::kMax], ee[USetEmul::kMax];
    for (int i = 0; i < cnt; ++i) {
      const int u = cand_u[q0 + i];
      loc[i] = pos[u];
      ee[i] = cand_e[q0 + i];
      firstq[u] = INT_MAX;
    }
    if (!record) continue;
    for (int i = 1; i < cnt; ++i) {
      const int l = loc[i], e = ee[i];
      int j = i - 1;
      while (j >= 0 && loc[j] > l) {
        loc[j + 1] = loc[j];
        ee[j + 1] = ee[j];
        --j;
      }
      loc[j + 1] = l;
      ee[j + 1] = e;
    }
    for (int i = 0; i < cnt; ++i) {
      ocol[q0 + i] = loc[i];
      oval[q0 + i] = P.val[ee[i]];
      erow[q0 + i] = t;
      if (last) cand_u[q0 + i] = gcolp[ee[i]];     // global column ids of the outermost hop (fused feature gather)
    }
  }
  __syncthreads();
  return n_rows + (int)n_new;
}

// ------------------------------------------------------------------------------------------ serial part (1 CTA)
__global__ void __launch_bounds__(kThreads)
sample_serial_kernel(Geom G, Ptrs P, int n_class, const int32_t* __restrict__ batch,
                     const int32_t* __restrict__ batch_off, const uint8_t* __restrict__ materialise) {
  __shared__ long long sm[kWarps + 1];
  __shared__ long long s_red;
  const int rem0 = (int)P.step_info[1], next0 = (int)P.step_info[2];
  long long draw_base = 0;
  const int k_last = G.fan[G.nh - 1];
  for (int c = 0; c < n_class; ++c) {
    const bool keep = materialise == nullptr || materialise[c] != 0;
    int32_t* nid = P.nid + (int64_t)c * G.lcap[G.nh];
    int32_t* pos = P.pos + (int64_t)c * G.n;
    int32_t* lc = P.level_count + (int64_t)c * (G.nh + 1);
    const int b0 = batch_off[c];
    int n_rows = batch_off[c + 1] - b0;
    for (int t = threadIdx.x; t < n_rows; t += blockDim.x) {
      const int v = batch[b0 + t];
      nid[t] = v;
      pos[v] = t;
    }
    if (threadIdx.x == 0) lc[0] = n_rows;
    __syncthreads();
    for (int h = 0; h + 1 < G.nh; ++h) {
      long long dr;
      n_rows = do_hop(G, P, c, h, keep, draw_base, rem0, next0, n_rows, &dr, sm);
      draw_base += dr;
      if (threadIdx.x == 0) lc[h + 1] = n_rows;
    }
    // segment of the stream that the class's last hop will consume
    long long mine = 0;
    for (int t = threadIdx.x; t < n_rows; t += blockDim.x) {
      const int v = nid[t];
      if (P.rowptr[v + 1] - P.rowptr[v] > k_last) mine += k_last;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (threadIdx.x == 0) s_red = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(reinterpret_cast<unsigned long long*>(&s_red), (unsigned long long)mine);
    __syncthreads();
    if (threadIdx.x == 0) P.last_off[c] = draw_base;
    draw_base += s_red;
    __syncthreads();
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
  }
}

// ------------------------------------------------------------------------------------------ last hop, 1 CTA / class
// transposed class block: rows = class-local columns, entries in source-row order (what a stable sort by column gives)
__device__ void build_transpose(const Geom& G, const Ptrs& P, int c, int h, int n_rows, int n_cols, int32_t* cursor_sm,
                                long long* sm) {
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
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int e0 = 0; e0 < nnz; e0 += 32) {
      const int e = e0 + lane;
      const bool live = e < nnz;
      const int cj = live ? ocol[e] : -1 - lane;
      const unsigned m = __match_any_sync(0xffffffffu, cj);
      const int rank = __popc(m & ((1u << lane) - 1u));
      int base = 0;
      if (live) base = cursor[cj];
      __syncwarp();
      if (live && rank == 0) cursor[cj] = base + __popc(m);
      __syncwarp();
      if (live) {
        tc[base + rank] = erow[e];
        tv[base + rank] = oval[e];
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
sample_last_kernel(Geom G, Ptrs P, int n_class, const uint8_t* __restrict__ materialise) {
  extern __shared__ int32_t cursor_sm[];
  __shared__ long long sm[kWarps + 1];
  const int c = blockIdx.x;
  if (c >= n_class) return;
  const bool keep = materialise == nullptr || materialise[c] != 0;
  const int rem0 = (int)P.step_info[1], next0 = (int)P.step_info[2];
  int32_t* nid = P.nid + (int64_t)c * G.lcap[G.nh];
  int32_t* pos = P.pos + (int64_t)c * G.n;
  int32_t* lc = P.level_count + (int64_t)c * (G.nh + 1);
  int n_rows = lc[G.nh - 1];
  if (keep) {
    long long dr;
    n_rows = do_hop(G, P, c, G.nh - 1, true, P.last_off[c], rem0, next0, n_rows, &dr, sm);
    if (threadIdx.x == 0) lc[G.nh] = n_rows;
    __syncthreads();
    for (int h = 0; h < G.nh; ++h) build_transpose(G, P, c, h, lc[h], lc[h + 1], cursor_sm, sm);
  }
  for (int t = threadIdx.x; t < n_rows; t += blockDim.x) pos[nid[t]] = -1;
}

// ------------------------------------------------------------------------------------------ packing
// Same layout and `desc` table as gs_sampler_finish_step (host_sampler.cpp); offsets are byte offsets into `out`.
__global__ void pack_offsets_kernel(Geom G, Ptrs P, int n_class, const uint8_t* __restrict__ materialise, int has_labels,
                                    int64_t out_cap, int64_t* __restrict__ desc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int nh = G.nh, al = G.align;
  for (int i = 0; i < 64; ++i) desc[i] = -1;
  for (int l = 0; l <= nh; ++l) {
    int32_t* s = P.seg + (int64_t)l * (n_class + 1);
    s[0] = 0;
    for (int c = 0; c < n_class; ++c) {
      const bool keep = materialise == nullptr || materialise[c] != 0;
      const int cnt = keep ? P.level_count[(int64_t)c * (nh + 1) + l] : 0;
      s[c + 1] = s[c] + (cnt + al - 1) / al * al;
    }
  }
  for (int h = 0; h < nh; ++h) {
    int64_t* e = P.eoff + (int64_t)h * (n_class + 1);
    e[0] = 0;
    for (int c = 0; c < n_class; ++c) {
      const bool keep = materialise == nullptr || materialise[c] != 0;
      const int rows = P.level_count[(int64_t)c * (nh + 1) + h];
      const int64_t m = keep ? (int64_t)(P.rowoff[h] + (int64_t)c * (G.lcap[h] + 1))[rows] : 0;
      e[c + 1] = e[c] + m;
    }
  }
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

__global__ void __launch_bounds__(256)
pack_copy_kernel(Geom G, Ptrs P, int n_class, const uint8_t* __restrict__ materialise, uint8_t* __restrict__ out,
                 const int64_t* __restrict__ desc) {
  const int c = blockIdx.x;
  const int nh = G.nh;
  if (desc[63] != 0) return;
  const bool keep = materialise == nullptr || materialise[c] != 0;
  const int tid = threadIdx.x, nt = blockDim.x;
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
  S->arena_bytes = off;
  cudaError_t e = cudaMalloc(&S->arena, (size_t)off);
  if (e != cudaSuccess) {
    set_error("gs_dsampler_create: cudaMalloc", e);
    delete S;
    return nullptr;
  }
  for (const Piece& p : pieces) *p.dst = static_cast<uint8_t*>(S->arena) + p.at;
  cudaStream_t st = as_stream(stream);
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
  if (cudaFuncSetAttribute(sample_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCursorSmemInts * 4) !=
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
  sample_serial_kernel<<<1, kThreads, 0, st>>>(G, S->P, n_class, d_batch, d_batch_off, d_materialise);
  if ((rc = finish_launch("dsampler_serial"))) return rc;
  sample_last_kernel<<<n_class, kThreads, kCursorSmemInts * 4, st>>>(G, S->P, n_class, d_materialise);
  if ((rc = finish_launch("dsampler_last_hop"))) return rc;
  pack_offsets_kernel<<<1, 32, 0, st>>>(G, S->P, n_class, d_materialise, S->has_labels ? 1 : 0, out_cap, d_desc);
  if ((rc = finish_launch("dsampler_pack_offsets"))) return rc;
  pack_copy_kernel<<<n_class, 256, 0, st>>>(G, S->P, n_class, d_materialise, d_out, d_desc);
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
