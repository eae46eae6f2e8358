// K1/K2/K3: CSR SpMM propagation over the real graph (full graph and sampled blocks).
//
//   Y[r,:] (+)= sum_e val[e] * X[col[e],:]
//
// Replaces torch_sparse.matmul at graphslim/models/sgc.py:47,51 and models/layers.py:41; with
// global column ids it also performs the features[n_id] gather of condensation/gcond_base.py:214.
//
// Mapping (HBM/L2-bound integer+fp32 gather work, no tensor cores):
//   * one warp per work item (a row, or a bounded slice of a long row for power-law tails);
//   * the warp stages 32 (col,val) pairs at a time in shared memory with one coalesced load,
//     then walks them; every lane owns NV float4 columns of the feature row, so a feature-row
//     gather is one fully coalesced 512 B * NV request per non-zero (rows are padded to 32 B);
//   * narrow widths (F/4 < 32, e.g. the class-width propagations of SGC) split the warp into
//     groups that take alternate non-zeros and are combined with warp shuffles.
//
// Two generations of the wide kernel live here (gs_spmm_set_tuning picks one; both produce the same bits:
// one accumulator per output element, non-zeros in CSR order, one fmaf per non-zero):
//   v1  one work item per warp, (col,val) staged through shared memory;
//   v2  a warp owns R consecutive work items and software-pipelines them: the item descriptors are fetched
//       with one coalesced load, the next 32 (col,val) pairs are always in flight in registers (across row
//       boundaries) while the current ones are consumed, and the feature-row gathers are issued UNR at a
//       time before their FMAs, so a warp keeps UNR*NV 512-byte requests outstanding instead of stalling
//       on the rowptr -> col -> X dependent chain once per row.  Outputs are written with streaming stores
//       and (col,val) are read evict-first so the gathered rows of X keep the L2.
#include <cstdlib>

#include "common.cuh"

namespace gs {

constexpr int kWarpsPerBlock = 8;

template <int VEC>
struct V;
template <>
struct V<4> {
  using T = float4;
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ T ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ T ldrw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, T v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ void stcs(float* p, T v) { __stcs(reinterpret_cast<float4*>(p), v); }
  static __device__ __forceinline__ void fma(T& a, float s, T x) {
    a.x = fmaf(s, x.x, a.x);
    a.y = fmaf(s, x.y, a.y);
    a.z = fmaf(s, x.z, a.z);
    a.w = fmaf(s, x.w, a.w);
  }
  static __device__ __forceinline__ T add(T a, T b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
  static __device__ __forceinline__ T shfl_xor(T a, int o) {
    return make_float4(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o),
                       __shfl_xor_sync(0xffffffffu, a.z, o), __shfl_xor_sync(0xffffffffu, a.w, o));
  }
  static __device__ __forceinline__ void atomic_add(float* p, T v) {
    atomicAdd(reinterpret_cast<float4*>(p), v);
  }
};
template <>
struct V<1> {
  using T = float;
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ T ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ T ldrw(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, T v) { *p = v; }
  static __device__ __forceinline__ void stcs(float* p, T v) { __stcs(p, v); }
  static __device__ __forceinline__ void fma(T& a, float s, T x) { a = fmaf(s, x, a); }
  static __device__ __forceinline__ T add(T a, T b) { return a + b; }
  static __device__ __forceinline__ T shfl_xor(T a, int o) { return __shfl_xor_sync(0xffffffffu, a, o); }
  static __device__ __forceinline__ void atomic_add(float* p, T v) { atomicAdd(p, v); }
};

// sequential streams (rowptr, col, val): pull 256 B into the L2 per miss so that the neighbouring warps' reads of the
// same stream are L2 hits instead of another DRAM round trip in front of their gathers
__device__ __forceinline__ int32_t ld_seq(const int32_t* p) {
  int32_t v;
  asm volatile("ld.global.nc.L2::256B.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_seq(const float* p) {
  float v;
  asm volatile("ld.global.nc.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// Work items: [0, n_rows) = one row each (direct store); [n_rows, n_rows + n_chunks) = slices of the long rows
// (degree > long_thr), accumulated with atomics into rows the caller has zeroed.  Row items skip long rows.
struct Items {
  int32_t n_items;
  int32_t n_rows;
  int32_t long_thr;
  const int32_t* rowptr;
  const int32_t* chunk_row;
  const int32_t* chunk_beg;
  const int32_t* chunk_end;
};

// returns false when the item has nothing to do; `atomic` tells the caller how to commit
__device__ __forceinline__ bool item_range(const Items& it, int i, int& row, int& beg, int& end, bool& atomic) {
  if (i >= it.n_rows) {
    const int c = i - it.n_rows;
    row = it.chunk_row[c];
    beg = it.chunk_beg[c];
    end = it.chunk_end[c];
    atomic = true;
    return true;
  }
  row = i;
  beg = it.rowptr[i];
  end = it.rowptr[i + 1];
  atomic = false;
  return !(it.chunk_row && end - beg > it.long_thr);
}

// Wide rows: L = ceil(F/VEC) >= 32 vector columns, tiled by 32*NV per warp (blockIdx.y = column tile).
template <int VEC, int NV, int WPB, int UNR>
__global__ void __launch_bounds__(WPB * 32)
spmm_wide_kernel(Items it, const int32_t* __restrict__ col, const float* __restrict__ val,
                 const float* __restrict__ X, int64_t ldx, int L, float* __restrict__ Y, int64_t ldy, int mode,
                 int flags) {
  using VT = typename V<VEC>::T;
  __shared__ int32_t s_col[WPB][32];
  __shared__ float s_val[WPB][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * WPB + warp;
  if (item >= it.n_items) return;
  int row, beg, end;
  bool atomic;
  if ((flags & 4) && item < it.n_rows) {
    row = item;
    beg = ld_seq(it.rowptr + item);
    end = ld_seq(it.rowptr + item + 1);
    atomic = false;
    if (it.chunk_row && end - beg > it.long_thr) return;
  } else if (!item_range(it, item, row, beg, end, atomic)) {
    return;
  }
  if (atomic) mode = 2;
  const int tile0 = blockIdx.y * (32 * NV);  // first vector column of this tile
  VT acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = V<VEC>::zero();
  bool live[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) live[v] = (tile0 + v * 32 + lane) < L;

  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    __syncwarp();
    if (lane < cnt) {
      if (flags & 4) {
        s_col[warp][lane] = ld_seq(col + base + lane);
        s_val[warp][lane] = ld_seq(val + base + lane);
      } else if (flags & 2) {
        s_col[warp][lane] = __ldcs(col + base + lane);
        s_val[warp][lane] = __ldcs(val + base + lane);
      } else {
        s_col[warp][lane] = __ldg(col + base + lane);
        s_val[warp][lane] = __ldg(val + base + lane);
      }
    }
    __syncwarp();
#pragma unroll UNR
    for (int j = 0; j < cnt; ++j) {
      const float a = s_val[warp][j];
      const float* xr = X + (int64_t)s_col[warp][j] * ldx + (int64_t)(tile0 + lane) * VEC;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (live[v]) V<VEC>::fma(acc[v], a, V<VEC>::ld(xr + (int64_t)v * 32 * VEC));
      }
    }
  }
  float* yr = Y + (int64_t)row * ldy + (int64_t)(tile0 + lane) * VEC;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (!live[v]) continue;
    float* p = yr + (int64_t)v * 32 * VEC;
    if (mode == 2) {
      V<VEC>::atomic_add(p, acc[v]);
    } else if (mode == 1) {
      V<VEC>::st(p, V<VEC>::add(V<VEC>::ldrw(p), acc[v]));
    } else if (flags & 1) {
      V<VEC>::stcs(p, acc[v]);
    } else {
      V<VEC>::st(p, acc[v]);
    }
  }
}

// Narrow rows: L < 32.  LPG (power of two >= L) lanes cover the width, 32/LPG groups split the non-zeros.
template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
spmm_narrow_kernel(Items it, const int32_t* __restrict__ col, const float* __restrict__ val,
                   const float* __restrict__ X, int64_t ldx, int L, int LPG, float* __restrict__ Y, int64_t ldy,
                   int mode) {
  using VT = typename V<VEC>::T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kWarpsPerBlock + warp;
  if (item >= it.n_items) return;
  int row, beg, end;
  bool atomic;
  if (!item_range(it, item, row, beg, end, atomic)) return;
  if (atomic) mode = 2;
  const int G = 32 / LPG;
  const int g = lane / LPG, l = lane % LPG;
  const bool live = l < L;
  VT acc = V<VEC>::zero();
  for (int e = beg + g; e < end; e += G) {
    const float a = __ldg(val + e);
    const int32_t c = __ldg(col + e);
    if (live) V<VEC>::fma(acc, a, V<VEC>::ld(X + (int64_t)c * ldx + (int64_t)l * VEC));
  }
  for (int o = LPG; o < 32; o <<= 1) acc = V<VEC>::add(acc, V<VEC>::shfl_xor(acc, o));
  if (g == 0 && live) {
    float* p = Y + (int64_t)row * ldy + (int64_t)l * VEC;
    if (mode == 2) {
      V<VEC>::atomic_add(p, acc);
    } else if (mode == 1) {
      V<VEC>::st(p, V<VEC>::add(V<VEC>::ldrw(p), acc));
    } else {
      V<VEC>::st(p, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// v2: pipelined multi-item warps (see the header comment).
struct Tuning {
  int impl;    // 1 | 2
  int unr;     // 0 = by kernel generation / width, else 2 | 4 | 8
  int group;   // v2: 0 = by problem size, else items per warp (1..32)
  int flags;   // bit0: streaming stores of Y, bit1: evict-first loads of (col,val), bit2 (v1): 256 B L2 prefetch on
               // the sequential streams (rowptr, col, val)
  int wpb;     // v1: warps per CTA (8 | 4 | 2)
  int max_nv;  // v1: widest column tile in units of 32 float4 (8 | 4 | 2 | 1); narrower tiles = fewer registers, more warps
};
// Measured on B200 (profiles/r1_spmm_sweep_v3.json): occupancy beats per-warp pipelining for this gather -- the
// 32-register v1 kernel with 64 resident warps per SM is 1.4-3.7x faster than v2 at every sweep point, so v1 is the
// default and v2 stays as the documented negative result behind the knob.
// Auto settings (0) of v1, from profiles/r1_spmm_sweep_v3.json: one column tile (F <= 128 floats) -> 4-warp CTAs with
// 8 gathers in flight (1.15-1.28x over 8 warps / 4 gathers); wider rows -> 2-warp CTAs, 4 gathers (1.05-1.12x).  Dense
// graphs with wide rows additionally gain 1.2x from 64-float4 column tiles (max_nv = 2, unr = 8), which the host
// wrapper selects from nnz / n_rows -- the C entry point does not know nnz without a device read.
static Tuning env_tuning() {
  Tuning t{1, 0, 0, 0, 0, 0};
  const char* e = getenv("GS_SPMM_IMPL");     // escape hatch for A/B runs
  if (e && atoi(e) == 2) t = Tuning{2, 0, 0, 3, 0, 0};
  return t;
}
static Tuning g_tune = env_tuning();

constexpr int kV2Warps = 4;   // small CTAs: the register-heavy instantiations still fill an SM in 4-warp steps

template <int VEC, int NV, int UNR>
__global__ void __launch_bounds__(kV2Warps * 32)
spmm_wide_v2_kernel(Items it, const int32_t* __restrict__ col, const float* __restrict__ val,
                    const float* __restrict__ X, int64_t ldx, int L, float* __restrict__ Y, int64_t ldy, int mode0,
                    int R, int flags) {
  using VT = typename V<VEC>::T;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * kV2Warps + (threadIdx.x >> 5);
  const int64_t g0 = wid * R;
  if (g0 >= it.n_items) return;
  const int nseg = (int)min((int64_t)R, (int64_t)it.n_items - g0);
  const bool cs_store = flags & 1, cs_load = flags & 2;

  // lane i holds the descriptor of the warp's i-th item
  int s_row = 0, s_beg = 0, s_end = 0;
  bool s_atomic = false, s_valid = false;
  if (lane < nseg) s_valid = item_range(it, (int)(g0 + lane), s_row, s_beg, s_end, s_atomic);
  if (!s_valid) s_end = s_beg;

  const int tile0 = blockIdx.y * (32 * NV);
  bool live[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) live[v] = (tile0 + v * 32 + lane) < L;
  const int64_t lane_off = (int64_t)(tile0 + lane) * VEC;

  auto load_pair = [&](int base, int end, int& c, float& a) {
    c = 0;
    a = 0.f;
    if (base + lane < end) {
      if (cs_load) {
        c = __ldcs(col + base + lane);
        a = __ldcs(val + base + lane);
      } else {
        c = __ldg(col + base + lane);
        a = __ldg(val + base + lane);
      }
    }
  };

  // cursor: item ci, 32-slice [cbase, min(cbase+32, cend)), its (col,val) in (c, a)
  int ci = 0;
  int cbase = __shfl_sync(FULL, s_beg, 0), cend = __shfl_sync(FULL, s_end, 0);
  int c;
  float a;
  load_pair(cbase, cend, c, a);
  VT acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = V<VEC>::zero();

  while (ci < nseg) {
    // where the next slice is (same item, or the head of the next item) -- and put its loads in flight
    const bool last_slice = cbase + 32 >= cend;
    int ni = ci, nbase = cbase + 32, nend = cend;
    if (last_slice) {
      ni = ci + 1;
      const int src = min(ni, 31);
      nbase = __shfl_sync(FULL, s_beg, src);
      nend = __shfl_sync(FULL, s_end, src);
      if (ni >= nseg) nend = nbase;
    }
    int c2;
    float a2;
    load_pair(nbase, nend, c2, a2);

    const int cnt = min(32, cend - cbase);
    for (int j = 0; j < cnt; j += UNR) {
      VT x[UNR][NV];
      float aj[UNR];
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const int src = min(j + k, 31);
        const int cj = __shfl_sync(FULL, c, src);
        aj[k] = __shfl_sync(FULL, a, src);
        const float* xr = X + (int64_t)cj * ldx + lane_off;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          x[k][v] = V<VEC>::zero();
          if (j + k < cnt && live[v]) x[k][v] = V<VEC>::ld(xr + (int64_t)v * 32 * VEC);
        }
      }
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        if (j + k < cnt) {
#pragma unroll
          for (int v = 0; v < NV; ++v) V<VEC>::fma(acc[v], aj[k], x[k][v]);
        }
      }
    }

    if (last_slice) {
      const bool valid = __shfl_sync(FULL, (int)s_valid, ci) != 0;
      const bool atomic = __shfl_sync(FULL, (int)s_atomic, ci) != 0;
      const int row = __shfl_sync(FULL, s_row, ci);
      if (valid) {
        const int mode = atomic ? 2 : mode0;
        float* yr = Y + (int64_t)row * ldy + lane_off;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (!live[v]) continue;
          float* p = yr + (int64_t)v * 32 * VEC;
          if (mode == 2) {
            V<VEC>::atomic_add(p, acc[v]);
          } else if (mode == 1) {
            V<VEC>::st(p, V<VEC>::add(V<VEC>::ldrw(p), acc[v]));
          } else if (cs_store) {
            V<VEC>::stcs(p, acc[v]);
          } else {
            V<VEC>::st(p, acc[v]);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = V<VEC>::zero();
    }
    ci = ni;
    cbase = nbase;
    cend = nend;
    c = c2;
    a = a2;
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = kNumSMs;
  }
  return n;
}

template <int VEC, int NV>
static void launch_wide_v2(const Items& it, const int32_t* col, const float* val, const float* X, int64_t ldx, int L,
                           float* Y, int64_t ldy, int mode, int n_tiles, cudaStream_t st) {
  // items per warp: enough warps for >= 4 waves of 32 resident warps per SM, at most 8 items (16 for huge inputs)
  int R = g_tune.group;
  if (R <= 0) {
    const int64_t warps_wanted = (int64_t)sm_count() * 32 * 4;
    R = (int)min((int64_t)8, max((int64_t)1, (int64_t)it.n_items / warps_wanted));
  }
  R = max(1, min(R, 32));
  const int64_t n_warps = ((int64_t)it.n_items + R - 1) / R;
  dim3 grid((unsigned)((n_warps + kV2Warps - 1) / kV2Warps), n_tiles);
  int unr = g_tune.unr > 0 ? g_tune.unr : (NV <= 2 ? 8 : 4);
  if (NV * unr > 24) unr = NV * 4 > 24 ? 2 : 4;          // keep the in-flight tile within the register file
#define GS_V2(U) spmm_wide_v2_kernel<VEC, NV, U><<<grid, kV2Warps * 32, 0, st>>>(it, col, val, X, ldx, L, Y, ldy, mode, R, g_tune.flags)
  if (unr >= 8) GS_V2(8);
  else if (unr >= 4) GS_V2(4);
  else GS_V2(2);
#undef GS_V2
}

template <int VEC>
static int launch_spmm(const Items& it, const int32_t* col, const float* val, const float* X, int64_t ldx, int F,
                       float* Y, int64_t ldy, int mode, cudaStream_t st) {
  const int L = (F + VEC - 1) / VEC;
  const int gx = (it.n_items + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (gx == 0) return GS_OK;
  if (L < 32) {
    int LPG = 1;
    while (LPG < L) LPG <<= 1;
    spmm_narrow_kernel<VEC><<<gx, kWarpsPerBlock * 32, 0, st>>>(it, col, val, X, ldx, L, LPG, Y, ldy, mode);
    return finish_launch("spmm_narrow");
  }
  const int max_nv = (g_tune.impl == 1 && g_tune.max_nv > 0) ? min(g_tune.max_nv, 8) : 8;
  const int tile_w = 32 * max_nv;                                                     // float4 columns per tile
  const int ntiles = (L + tile_w - 1) / tile_w;
  const int per_tile = (L + ntiles - 1) / ntiles;
  const int nv = (per_tile + 31) / 32;
  if (g_tune.impl == 2) {
    const int n_tiles = (L + 32 * nv - 1) / (32 * nv);
    switch (nv) {
#define GS_V2_CASE(NVV) case NVV: launch_wide_v2<VEC, NVV>(it, col, val, X, ldx, L, Y, ldy, mode, n_tiles, st); break;
      GS_V2_CASE(1) GS_V2_CASE(2) GS_V2_CASE(3) GS_V2_CASE(4) GS_V2_CASE(5) GS_V2_CASE(6) GS_V2_CASE(7) GS_V2_CASE(8)
#undef GS_V2_CASE
      default:
        return GS_EINVAL;
    }
    return finish_launch("spmm_wide_v2");
  }
  const int wpb = g_tune.wpb > 0 ? g_tune.wpb : (nv == 1 ? 4 : 2);
  const int unr = g_tune.unr > 0 ? (g_tune.unr == 8 ? 8 : 4) : (nv == 1 ? 8 : 4);
  const int gxw = (it.n_items + wpb - 1) / wpb;
  dim3 grid(gxw, (L + 32 * nv - 1) / (32 * nv));
#define GS_SPMM_LAUNCH(NVV, W, U) \
  spmm_wide_kernel<VEC, NVV, W, U><<<grid, W * 32, 0, st>>>(it, col, val, X, ldx, L, Y, ldy, mode, g_tune.flags)
#define GS_SPMM_CASE(NVV)                                              \
  case NVV:                                                            \
    if (unr == 8) {                                                    \
      if (wpb == 8) GS_SPMM_LAUNCH(NVV, 8, 8);                         \
      else if (wpb == 4) GS_SPMM_LAUNCH(NVV, 4, 8);                    \
      else GS_SPMM_LAUNCH(NVV, 2, 8);                                  \
    } else {                                                           \
      if (wpb == 8) GS_SPMM_LAUNCH(NVV, 8, 4);                         \
      else if (wpb == 4) GS_SPMM_LAUNCH(NVV, 4, 4);                    \
      else GS_SPMM_LAUNCH(NVV, 2, 4);                                  \
    }                                                                  \
    break;
  switch (nv) {
    GS_SPMM_CASE(1)
    GS_SPMM_CASE(2)
    GS_SPMM_CASE(3)
    GS_SPMM_CASE(4)
    GS_SPMM_CASE(5)
    GS_SPMM_CASE(6)
    GS_SPMM_CASE(7)
    GS_SPMM_CASE(8)
    default:
      return GS_EINVAL;
  }
#undef GS_SPMM_CASE
#undef GS_SPMM_LAUNCH
  return finish_launch("spmm_wide");
}

static inline bool vec4_ok(const float* X, int64_t ldx, const float* Y, int64_t ldy, int F) {
  return (F % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
         ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
}

// ------------------------------------------------------------------------------------------------
// scatter form (transpose-free backward through a rectangular block)
template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
spmm_scatter_kernel(int n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const float* __restrict__ val, const float* __restrict__ dY, int64_t ldy, int L,
                    float* __restrict__ dX, int64_t ldx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + warp;
  if (row >= n_rows) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  for (int l = lane; l < L; l += 32) {
    const typename V<VEC>::T g = V<VEC>::ld(dY + (int64_t)row * ldy + (int64_t)l * VEC);
    for (int e = beg; e < end; ++e) {
      typename V<VEC>::T c = V<VEC>::zero();
      V<VEC>::fma(c, __ldg(val + e), g);
      V<VEC>::atomic_add(dX + (int64_t)__ldg(col + e) * ldx + (int64_t)l * VEC, c);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gather_rows_kernel(int n, const int32_t* __restrict__ idx, const float* __restrict__ X, int64_t ldx, int L,
                   float* __restrict__ out, int64_t ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + warp;
  if (r >= n) return;
  const float* src = X + (int64_t)__ldg(idx + r) * ldx;
  float* dst = out + (int64_t)r * ldo;
  for (int l = lane; l < L; l += 32) V<VEC>::st(dst + (int64_t)l * VEC, V<VEC>::ld(src + (int64_t)l * VEC));
}

// zero the output rows of the long rows before their slices are accumulated
template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
zero_long_rows_kernel(int n_chunks, const int32_t* __restrict__ chunk_row, const int32_t* __restrict__ chunk_beg,
                      const int32_t* __restrict__ rowptr, int L, float* __restrict__ Y, int64_t ldy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kWarpsPerBlock + warp;
  if (c >= n_chunks) return;
  const int row = chunk_row[c];
  if (chunk_beg[c] != rowptr[row]) return;        // only the first slice of a row clears it
  for (int l = lane; l < L; l += 32) V<VEC>::st(Y + (int64_t)row * ldy + (int64_t)l * VEC, V<VEC>::zero());
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
csr_gcn_norm_kernel(int n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const float* __restrict__ a, const double* __restrict__ r, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + warp;
  if (row >= n_rows) return;
  const double ri = r[row];
  for (int e = rowptr[row] + lane; e < rowptr[row + 1]; e += 32) {
    // (D^-1/2 A) D^-1/2 in float64, one rounding to fp32: graphslim/utils.py:451-458 then :664
    const double left = __dmul_rn(ri, (double)a[e]);
    out[e] = (float)__dmul_rn(left, r[col[e]]);
  }
}

}  // namespace gs

extern "C" {

int gs_spmm_csr_f32(int32_t n_rows, const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                    int64_t ldx, int32_t F, float* Y, int64_t ldy, int accumulate, int32_t n_chunks,
                    int32_t long_thr, const int32_t* chunk_row, const int32_t* chunk_beg, const int32_t* chunk_end,
                    void* stream) {
  if (n_rows == 0) return GS_OK;
  GS_REQUIRE(n_rows >= 0 && F > 0 && rowptr && X && Y && ldx >= F && ldy >= F);
  GS_REQUIRE(n_chunks == 0 || (chunk_row && chunk_beg && chunk_end && long_thr > 0));
  gs::Items it{n_rows + n_chunks, n_rows, long_thr, rowptr, n_chunks > 0 ? chunk_row : nullptr, chunk_beg, chunk_end};
  const int mode = accumulate ? 1 : 0;
  cudaStream_t st = gs::as_stream(stream);
  const bool v4 = gs::vec4_ok(X, ldx, Y, ldy, F);
  if (n_chunks > 0 && !accumulate) {
    const int gx = (n_chunks + gs::kWarpsPerBlock - 1) / gs::kWarpsPerBlock;
    if (v4) gs::zero_long_rows_kernel<4><<<gx, gs::kWarpsPerBlock * 32, 0, st>>>(n_chunks, chunk_row, chunk_beg, rowptr, F / 4, Y, ldy);
    else gs::zero_long_rows_kernel<1><<<gx, gs::kWarpsPerBlock * 32, 0, st>>>(n_chunks, chunk_row, chunk_beg, rowptr, F, Y, ldy);
    const int rc = gs::finish_launch("zero_long_rows");
    if (rc) return rc;
  }
  if (v4) return gs::launch_spmm<4>(it, col, val, X, ldx, F, Y, ldy, mode, st);
  return gs::launch_spmm<1>(it, col, val, X, ldx, F, Y, ldy, mode, st);
}

// L2-resident column tiling for X larger than the L2 (Reddit shape: 233 K rows x 602 floats = 561 MB against 126 MB):
// the product is run tile_cols columns at a time, so the gathered slice X[:, c0:c0+tile_cols] (n_src x tile_cols x 4 B,
// chosen by the caller to fit the L2) stays resident while every row sweeps it; each pass re-reads the 8 B / nnz of
// (col, val) but the 4 F B / nnz of gathered feature rows come from the L2 instead of HBM.  Full slices of >= 128 floats give
// the bits of the untiled call (every output element accumulates its non-zeros in CSR order); narrower slices (and a
// narrower remainder) run the split-warp kernel and agree to fp32 reassociation.
int gs_spmm_csr_tiled_f32(int32_t n_rows, const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                          int64_t ldx, int32_t F, float* Y, int64_t ldy, int accumulate, int32_t n_chunks,
                          int32_t long_thr, const int32_t* chunk_row, const int32_t* chunk_beg,
                          const int32_t* chunk_end, int32_t tile_cols, void* stream) {
  GS_REQUIRE(tile_cols > 0 && tile_cols % 4 == 0);
  for (int32_t c0 = 0; c0 < F; c0 += tile_cols) {
    const int32_t w = F - c0 < tile_cols ? F - c0 : tile_cols;
    const int rc = gs_spmm_csr_f32(n_rows, rowptr, col, val, X + c0, ldx, w, Y + c0, ldy, accumulate, n_chunks, long_thr,
                                   chunk_row, chunk_beg, chunk_end, stream);
    if (rc) return rc;
  }
  return GS_OK;
}

int gs_spmm_set_tuning(int impl, int unr, int group, int flags, int wpb, int max_nv) {
  GS_REQUIRE((impl == 1 || impl == 2) && (unr == 0 || unr == 2 || unr == 4 || unr == 8) && group >= 0 && group <= 32 &&
             flags >= 0 && flags <= 7 && (wpb == 0 || wpb == 8 || wpb == 4 || wpb == 2) &&
             (max_nv == 0 || max_nv == 8 || max_nv == 4 || max_nv == 2 || max_nv == 1));
  gs::g_tune = gs::Tuning{impl, unr, group, flags, wpb, max_nv};
  return GS_OK;
}

int gs_spmm_csr_scatter_f32(int32_t n_rows, const int32_t* rowptr, const int32_t* col, const float* val,
                            const float* dY, int64_t ldy, int32_t F, float* dX, int64_t ldx, void* stream) {
  if (n_rows == 0) return GS_OK;
  GS_REQUIRE(n_rows >= 0 && F > 0 && rowptr && dY && dX && ldx >= F && ldy >= F);
  if (n_rows == 0) return GS_OK;
  const int gx = (n_rows + gs::kWarpsPerBlock - 1) / gs::kWarpsPerBlock;
  cudaStream_t st = gs::as_stream(stream);
  if (gs::vec4_ok(dY, ldy, dX, ldx, F)) {
    gs::spmm_scatter_kernel<4><<<gx, gs::kWarpsPerBlock * 32, 0, st>>>(n_rows, rowptr, col, val, dY, ldy, F / 4, dX, ldx);
  } else {
    gs::spmm_scatter_kernel<1><<<gx, gs::kWarpsPerBlock * 32, 0, st>>>(n_rows, rowptr, col, val, dY, ldy, F, dX, ldx);
  }
  return gs::finish_launch("spmm_scatter");
}

int gs_gather_rows_f32(int32_t n, const int32_t* idx, const float* X, int64_t ldx, int32_t F, float* out, int64_t ldo,
                       void* stream) {
  if (n == 0) return GS_OK;
  GS_REQUIRE(n >= 0 && F > 0 && idx && X && out && ldx >= F && ldo >= F);
  if (n == 0) return GS_OK;
  const int gx = (n + gs::kWarpsPerBlock - 1) / gs::kWarpsPerBlock;
  cudaStream_t st = gs::as_stream(stream);
  if (gs::vec4_ok(X, ldx, out, ldo, F)) {
    gs::gather_rows_kernel<4><<<gx, gs::kWarpsPerBlock * 32, 0, st>>>(n, idx, X, ldx, F / 4, out, ldo);
  } else {
    gs::gather_rows_kernel<1><<<gx, gs::kWarpsPerBlock * 32, 0, st>>>(n, idx, X, ldx, F, out, ldo);
  }
  return gs::finish_launch("gather_rows");
}

int gs_csr_gcn_norm_f64(int32_t n_rows, const int32_t* rowptr, const int32_t* col, const float* a, const double* r,
                        float* val_out, void* stream) {
  GS_REQUIRE(n_rows >= 0 && rowptr && col && a && r && val_out);
  if (n_rows == 0) return GS_OK;
  const int gx = (n_rows + gs::kWarpsPerBlock - 1) / gs::kWarpsPerBlock;
  gs::csr_gcn_norm_kernel<<<gx, gs::kWarpsPerBlock * 32, 0, gs::as_stream(stream)>>>(n_rows, rowptr, col, a, r, val_out);
  return gs::finish_launch("csr_gcn_norm");
}

}  // extern "C"
