// Dense contractions on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   C[M,N] = alpha * op(A) * op(B) + beta * C      fp32 in HBM, BF16 multiplicands, fp32 accumulation in TMEM
//
// The synthetic side of GCond is a chain of dense products whose operands are fp32 (PGE's N'^2 x h x h
// layer, its two backward products, the class-column products of the condense model).  The north star asks
// for fp32-class agreement with the reference (1e-4), which a single BF16 (8-bit mantissa) product does not
// give, so every fp32 operand is split on the fly into BF16 "hi" and "lo" planes,
//      x = hi + lo + O(2^-17 |x|),   hi = bf16(x), lo = bf16(x - hi),
// and the product is accumulated as  hi*hi + hi*lo + lo*hi  (precision 1, ~2^-16 relative, three MMAs per
// k-step) or as hi*hi only (precision 2, one MMA, the stated looser bound).
//
// Operand B (the reused operand: weights, or the class-column matrix every M tile needs) is split once per call by
// `pack_b_kernel` into a BF16 hi/lo *tile image* in HBM that already has the shared-memory layout, so a stage of B
// is one contiguous bulk-async copy (cp.async.bulk -> SASS UBLKCP, the TMA engine's 1-D path) that completes on the
// stage's mbarrier with complete_tx; no tensor map and no conversion work in the main loop.  Operand A (streamed once)
// is converted on load by the producer warps.
//
// Structure (one persistent CTA per SM, 9 warps):
//   warps 0-3  producers  : coalesced float4 loads of the fp32 A tiles, split to BF16, stores into shared memory in
//                           the canonical K-major SWIZZLE_128B operand layout (8 rows x 128 B atoms, 16 B chunk
//                           index XOR row%8) -- also for a transposed source, so NN / NT / TN / TT all feed the same
//                           K-major descriptors; `fence.proxy.async` + mbarrier hand the stage to the MMA warp;
//                           producer thread 0 also issues the bulk copies of the B image (expect_tx on the barrier)
//   warps 4-7  epilogue   : tcgen05.ld of the 128 x BN fp32 accumulator (warp q owns TMEM lanes 32q..32q+31),
//                           transposed through a padded smem staging tile so every global store is a coalesced 128 B
//                           row segment; alpha/beta (or atomics for split-K)
//   warp  8    MMA issuer : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) and
//                           tcgen05.commit to free smem stages / publish accumulators; owns TMEM alloc/dealloc
// Accumulators are double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile
// i+1; K can be split across CTAs (dW = dY^T H with K = N'^2 has only h*h outputs).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace gs {
namespace tc {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BK = 64;         // 64 bf16 = one 128-byte swizzle row
constexpr int kProducerThreads = 256;                      // 8 warps: the A-operand conversion is latency bound
constexpr int kProducerWarps = kProducerThreads / 32;
constexpr int kEpilogueThreads = 128;
constexpr int kMmaWarp = kProducerWarps + kEpilogueThreads / 32;
constexpr int kThreads = kProducerThreads + kEpilogueThreads + 32;
constexpr int kStages = 2;

struct Params {
  int ta, tb;
  int M, N, K;
  float alpha, beta;
  const float* A;
  int64_t lda;
  const float* B;
  int64_t ldb;
  const uint8_t* Bimg;   // packed BF16 tile image of op(B)^T: [n_tile][k_block][plane][BN rows x 128 B]
  float* C;
  int64_t ldc;
  int splits;      // K split count (>1: epilogue uses atomics, C pre-scaled by the caller)
  int kchunk;      // K range per split (multiple of BK)
  int tiles_m, tiles_n;
  // grouped mode (seg != nullptr): group g contracts rows seg[g]..seg[g+1] (64-aligned) of A (K x M) and B (K x N)
  // into the column block out_block[g] of C; every group is split `splits` ways along K; C is pre-zeroed, atomics.
  const int32_t* seg;
  const int32_t* out_block;
  int groups;
};

// fused epilogue (never with split-K / grouped accumulation): v += bias[col]; relu; v = mask[row,col] > 0 ? v : 0.
// A separate kernel parameter: growing `Params` itself changes ptxas' register allocation of the producer loop (8 ->
// 24 bytes of spills and a 25 % slower streaming product were measured when these fields lived in Params).
struct Epi {
  const float* bias;
  int relu;
  const float* mask;
  int64_t ldmask;
};

struct Tile {
  int mb, nb, k_beg, k_end;
  int64_t c_col;   // column offset of this tile's output block in C
};

// Work item -> tile coordinates; false when the item is empty (all three warp roles skip it identically).
__device__ __forceinline__ bool decode_tile(const Params& p, int tile, Tile& t) {
  const int ks = tile % p.splits;
  int mn = tile / p.splits;
  if (p.seg) {
    const int per_group = p.tiles_m * p.tiles_n;
    const int g = mn / per_group;
    mn -= g * per_group;
    const int kb = __ldg(p.seg + g), ke = __ldg(p.seg + g + 1);
    const int nkb = (ke - kb + BK - 1) / BK;
    const int chunk = ((nkb + p.splits - 1) / p.splits) * BK;
    t.k_beg = kb + ks * chunk;
    t.k_end = min(ke, t.k_beg + chunk);
    t.c_col = (int64_t)__ldg(p.out_block + g) * p.N;
  } else {
    t.k_beg = ks * p.kchunk;
    t.k_end = min(p.K, t.k_beg + p.kchunk);
    t.c_col = 0;
  }
  t.mb = mn % p.tiles_m;
  t.nb = mn / p.tiles_m;
  return t.k_beg < t.k_end;
}
// first non-empty work item at or after `tile` (stride gridDim.x); false when the list is exhausted
__device__ __forceinline__ bool seek_tile(const Params& p, int& tile, int num_tiles, Tile& t) {
  while (tile < num_tiles) {
    if (decode_tile(p, tile, t)) return true;
    tile += gridDim.x;
  }
  return false;
}

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared, completion (bytes) signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t holder_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (ignored for swizzled K-major, set to 1),
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups), [46,48) version = 1, [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format BF16 (bits 7, 10),
// K-major A and B (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int umma_m, int umma_n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(umma_m >> 4) << 24);
}

// byte offset of element (row, k) [k in 0..63] inside one K-major SW128 tile of bf16 (rows x 64)
__device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  const int chunk = (k >> 3) ^ (row & 7);
  return (uint32_t)(row * 128 + chunk * 16 + (k & 7) * 2);
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// two floats -> packed bf16x2 hi word (element a in the low half) and packed lo word; one cvt.rn.bf16x2.f32 per word
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ------------------------------------------------------------------------------------------ producers
// A (ROWS x 64) fp32 tile is fetched into registers first (all loads in flight at once: the producers are latency
// bound otherwise) and converted / stored to shared memory later, so the fetch of k-block i+1 overlaps the wait for
// a free stage and the conversion of k-block i.
//
// Thread -> element mapping, ROWS*16 float4 per tile, kProducerThreads threads, NLD = ROWS*16/threads float4 each:
//   K-contiguous source  (element (r,k) at P[r*ld + k]): float4 f = tid + 128*i covers row f>>4, k = 4*(f&15)..+3
//   row-contiguous source (element (r,k) at P[k*ld + r]): warp w owns steps s = w + 4*i; a step covers 32 rows
//       (8 lanes x float4) x 4 k-pairs; each lane packs (k, k+1) into one bf16x2 word per row and rotates its four row
//       stores so the 32 lanes hit 32 distinct banks.
template <int ROWS>
struct TileRegs {
  static constexpr int NLD = ROWS * (BK / 4) / kProducerThreads;
  static_assert(NLD >= 2 && NLD % 2 == 0, "tile too small for the producer thread count");
  float4 v[NLD];
};

template <int ROWS>
__device__ __forceinline__ void fetch_k_contig(TileRegs<ROWS>& t, const float* __restrict__ P, int64_t ld,
                                               int n_rows_total, int k_end, int r0, int k0, int tid) {
  const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(P) & 15) == 0);
  if (vec_ok && r0 + ROWS <= n_rows_total && k0 + BK <= k_end) {
    // interior tile: no guards, one base pointer, constant strides (rows advance by kProducerThreads/16 per step)
    const float* src = P + (int64_t)(r0 + (tid >> 4)) * ld + k0 + (tid & 15) * 4;
    const int64_t step = (int64_t)(kProducerThreads / 16) * ld;
#pragma unroll
    for (int i = 0; i < TileRegs<ROWS>::NLD; ++i) t.v[i] = __ldg(reinterpret_cast<const float4*>(src + i * step));
    return;
  }
#pragma unroll
  for (int i = 0; i < TileRegs<ROWS>::NLD; ++i) {
    const int f = tid + kProducerThreads * i;
    const int r = f >> 4, c4 = f & 15;
    const int gr = r0 + r, gk = k0 + c4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < n_rows_total) {
      const float* src = P + (int64_t)gr * ld + gk;
      if (vec_ok && gk + 3 < k_end) {
        v = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        if (gk + 0 < k_end) v.x = __ldg(src + 0);
        if (gk + 1 < k_end) v.y = __ldg(src + 1);
        if (gk + 2 < k_end) v.z = __ldg(src + 2);
        if (gk + 3 < k_end) v.w = __ldg(src + 3);
      }
    }
    t.v[i] = v;
  }
}

template <int ROWS, bool kWithLo>
__device__ __forceinline__ void store_k_contig(const TileRegs<ROWS>& t, uint8_t* s_hi, uint8_t* s_lo, int tid) {
#pragma unroll
  for (int i = 0; i < TileRegs<ROWS>::NLD; ++i) {
    const int f = tid + kProducerThreads * i;
    const int r = f >> 4, c4 = f & 15;
    const float4 v = t.v[i];
    uint2 ph, pl;
    split_bf16x2(v.x, v.y, ph.x, pl.x);
    split_bf16x2(v.z, v.w, ph.y, pl.y);
    const uint32_t off = sw128_offset(r, c4 * 4);
    *reinterpret_cast<uint2*>(s_hi + off) = ph;
    if (kWithLo) *reinterpret_cast<uint2*>(s_lo + off) = pl;
  }
}

__device__ __forceinline__ float pick4(const float4& v, int j) {
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}

template <int ROWS>
__device__ __forceinline__ void fetch_r_contig(TileRegs<ROWS>& t, const float* __restrict__ P, int64_t ld,
                                               int n_rows_total, int k_end, int r0, int k0, int tid) {
  const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(P) & 15) == 0);
  const int warp = tid >> 5, lane = tid & 31;
  const int ri = lane & 7, kpi = lane >> 3;
#pragma unroll
  for (int i = 0; i < TileRegs<ROWS>::NLD / 2; ++i) {
    const int s = warp + (kProducerThreads / 32) * i;       // step: (row group of 32) x (k group of 8)
    const int rg = s / (BK / 8), kg = s % (BK / 8);
    const int gr = r0 + rg * 32 + ri * 4;
    const int gk = k0 + (kg * 4 + kpi) * 2;
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (vec_ok && gr + 3 < n_rows_total) {
      if (gk < k_end) va = __ldg(reinterpret_cast<const float4*>(P + (int64_t)gk * ld + gr));
      if (gk + 1 < k_end) vb = __ldg(reinterpret_cast<const float4*>(P + (int64_t)(gk + 1) * ld + gr));
    } else {
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (gr + j < n_rows_total) {
          if (gk < k_end) a[j] = __ldg(P + (int64_t)gk * ld + gr + j);
          if (gk + 1 < k_end) b[j] = __ldg(P + (int64_t)(gk + 1) * ld + gr + j);
        }
      }
      va = make_float4(a[0], a[1], a[2], a[3]);
      vb = make_float4(b[0], b[1], b[2], b[3]);
    }
    t.v[2 * i] = va;
    t.v[2 * i + 1] = vb;
  }
}

template <int ROWS, bool kWithLo>
__device__ __forceinline__ void store_r_contig(const TileRegs<ROWS>& t, uint8_t* s_hi, uint8_t* s_lo, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const int ri = lane & 7, kpi = lane >> 3;
#pragma unroll
  for (int i = 0; i < TileRegs<ROWS>::NLD / 2; ++i) {
    const int s = warp + (kProducerThreads / 32) * i;
    const int rg = s / (BK / 8), kg = s % (BK / 8);
    const int r4 = rg * 32 + ri * 4;
    const int kp = kg * 4 + kpi;
    const float4 va = t.v[2 * i], vb = t.v[2 * i + 1];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = (j + (ri >> 1)) & 3;
      const int row = r4 + jj;
      uint32_t ph, pl;
      split_bf16x2(pick4(va, jj), pick4(vb, jj), ph, pl);
      const uint32_t off = sw128_offset(row, kp * 2);
      *reinterpret_cast<uint32_t*>(s_hi + off) = ph;
      if (kWithLo) *reinterpret_cast<uint32_t*>(s_lo + off) = pl;
    }
  }
}

// ------------------------------------------------------------------------------------------ B tile image
// One CTA per (n_tile, k_block): the same split + swizzled placement the producers do, written to HBM.
template <int BN, bool kWithLo>
__global__ void __launch_bounds__(kProducerThreads) pack_b_kernel(const float* __restrict__ B, int64_t ldb, int tb, int N,
                                                                  int K, int kblocks, uint8_t* __restrict__ img) {
  constexpr int kPlanes = kWithLo ? 2 : 1;
  constexpr uint32_t kBBytes = BN * BK * 2;
  constexpr int SUB = BN < 128 ? BN : 128;              // rows converted per pass
  extern __shared__ __align__(16) uint8_t tile[];
  const int nb = blockIdx.x / kblocks, kb = blockIdx.x % kblocks;
#pragma unroll 1
  for (int r0 = 0; r0 < BN; r0 += SUB) {
    TileRegs<SUB> t;
    uint8_t* hi = tile + r0 * 128;
    uint8_t* lo = hi + kBBytes;
    if (tb) {
      fetch_k_contig<SUB>(t, B, ldb, N, K, nb * BN + r0, kb * BK, threadIdx.x);
      store_k_contig<SUB, kWithLo>(t, hi, lo, threadIdx.x);
    } else {
      fetch_r_contig<SUB>(t, B, ldb, N, K, nb * BN + r0, kb * BK, threadIdx.x);
      store_r_contig<SUB, kWithLo>(t, hi, lo, threadIdx.x);
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(img + (size_t)blockIdx.x * kPlanes * kBBytes);
  const uint4* src = reinterpret_cast<const uint4*>(tile);
  for (int i = threadIdx.x; i < (int)(kPlanes * kBBytes / 16); i += kProducerThreads) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------ kernel
// BN: accumulator width (columns of the output tile), multiple of 32, <= 256.  NPASS: 1 or 3.
// EPI: the fused bias / ReLU / mask epilogue is compiled in (a separate instantiation: the plain kernel keeps the
// leaner epilogue and register allocation the streaming PGE products were tuned with).
template <int BN, int NPASS, bool EPI>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(Params p, Epi ep) {
  constexpr bool kWithLo = NPASS == 3;
  constexpr int kPlanes = kWithLo ? 2 : 1;
  constexpr uint32_t kABytes = BM * BK * 2;      // one bf16 plane of the A tile
  constexpr uint32_t kBBytes = BN * BK * 2;
  constexpr uint32_t kStageBytes = kPlanes * (kABytes + kBBytes);
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  constexpr uint32_t kIdesc = make_idesc(BM, BN);

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the shared-space pointer (an integer round-trip would hand the
  // compiler a generic pointer: LD/ST instead of LDS/STS in the producers and the epilogue, and possible aliasing
  // between the staging tile and the global stores that serialises the epilogue's loads)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* stage_out = reinterpret_cast<float*>(smem + kStages * kStageBytes);   // 4 warps x 32 x 36 floats
  __shared__ __align__(8) uint64_t bar_full[kStages], bar_empty[kStages], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n * p.splits * (p.seg ? p.groups : 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), kProducerThreads + 1);   // +1: the expect_tx arrival for the B bulk copies
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_tfull[a]), 1);
      mbar_init(smem_u32(&bar_tempty[a]), kEpilogueThreads);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(&tmem_holder), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp < kProducerWarps) {
    // ============================== producers ==============================
    const int tid = threadIdx.x;
    const int kblocks_total = (p.K + BK - 1) / BK;
    // flattened (tile, k-block) work list; the fp32 A tiles of the next TWO items are kept in flight in registers
    struct Cursor {
      int tile, k0, k_end, mb, nb;
      bool valid;
    };
    auto start = [&](Cursor& c, int first_tile) {
      Tile tl;
      c.tile = first_tile;
      c.valid = seek_tile(p, c.tile, num_tiles, tl);
      if (c.valid) {
        c.mb = tl.mb;
        c.nb = tl.nb;
        c.k0 = tl.k_beg;
        c.k_end = tl.k_end;
      }
    };
    auto advance = [&](Cursor& c) {
      c.k0 += BK;
      if (c.k0 >= c.k_end) start(c, c.tile + gridDim.x);
    };
    auto fetch = [&](TileRegs<BM>& t, const Cursor& c) {
      if (!c.valid) return;
      if (p.ta) fetch_r_contig<BM>(t, p.A, p.lda, p.M, c.k_end, c.mb * BM, c.k0, tid);
      else fetch_k_contig<BM>(t, p.A, p.lda, p.M, c.k_end, c.mb * BM, c.k0, tid);
    };
    Cursor cur, ahead;
    start(cur, blockIdx.x);
    TileRegs<BM> r0, r1;
    fetch(r0, cur);
    ahead = cur;
    if (ahead.valid) advance(ahead);
    fetch(r1, ahead);
    uint32_t it = 0;
    while (cur.valid) {
      const TileRegs<BM> now = r0;
      const int c_k0 = cur.k0, c_nb = cur.nb;
      // rotate: the item after next starts loading before this one is converted
      r0 = r1;
      cur = ahead;
      if (ahead.valid) advance(ahead);
      fetch(r1, ahead);
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
      uint8_t* sa_hi = smem + s * kStageBytes;
      uint8_t* sb_hi = sa_hi + kPlanes * kABytes;
      uint8_t* sa_lo = sa_hi + kABytes;
      if (tid == 0) {   // B stage: one contiguous image block (hi and lo planes adjacent), DMA'd by the TMA engine
        const uint32_t bar = smem_u32(&bar_full[s]);
        const uint8_t* src = p.Bimg + ((size_t)c_nb * kblocks_total + (size_t)(c_k0 / BK)) * (kPlanes * kBBytes);
        mbar_arrive_expect_tx(bar, kPlanes * kBBytes);
        bulk_g2s(smem_u32(sb_hi), src, kPlanes * kBBytes, bar);
      }
      if (p.ta) store_r_contig<BM, kWithLo>(now, sa_hi, sa_lo, tid);
      else store_k_contig<BM, kWithLo>(now, sa_hi, sa_lo, tid);
      fence_proxy_async();               // generic-proxy stores -> visible to the tensor-core (async) proxy
      mbar_arrive(smem_u32(&bar_full[s]));
      ++it;
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ==============================
    uint32_t it = 0, tcount = 0;
    Tile tl;
    for (int tile = blockIdx.x; seek_tile(p, tile, num_tiles, tl); tile += gridDim.x, ++tcount) {
      const int k_beg = tl.k_beg, k_end = tl.k_end;
      const int acc = tcount & 1;
      const uint32_t acc_ph = (tcount >> 1) & 1;
      mbar_wait(smem_u32(&bar_tempty[acc]), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      bool first = true;
      for (int k0 = k_beg; k0 < k_end; k0 += BK, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa_hi = smem_u32(smem + s * kStageBytes);
          const uint32_t sb_hi = sa_hi + kPlanes * kABytes;
          const uint32_t sa_lo = sa_hi + kABytes;
          const uint32_t sb_lo = sb_hi + kBBytes;
#pragma unroll
          for (int pass = 0; pass < NPASS; ++pass) {
            const uint32_t a_base = (pass == 2) ? sa_lo : sa_hi;
            const uint32_t b_base = (pass == 1) ? sb_lo : sb_hi;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t ad = make_desc_k_sw128(a_base + kk * 32);
              const uint64_t bd = make_desc_k_sw128(b_base + kk * 32);
              umma_bf16(tmem_d, ad, bd, kIdesc, first ? 0u : 1u);
              first = false;
            }
          }
          umma_commit(smem_u32(&bar_empty[s]));          // frees the stage when these MMAs have read it
        }
        __syncwarp();
      }
      if (lane == 0) umma_commit(smem_u32(&bar_tfull[acc]));  // accumulator complete
      __syncwarp();
    }
  } else {
    // ============================== epilogue ==============================
    const int q = warp & 3;                                 // TMEM lane quarter owned by this warp
    uint32_t tcount = 0;
    Tile tl;
    const bool use_atomics = p.splits > 1 || p.seg != nullptr;
    for (int tile = blockIdx.x; seek_tile(p, tile, num_tiles, tl); tile += gridDim.x, ++tcount) {
      const int mb = tl.mb, nb = tl.nb;
      const int acc = tcount & 1;
      const uint32_t acc_ph = (tcount >> 1) & 1;
      mbar_wait(smem_u32(&bar_tfull[acc]), acc_ph);
      tc_fence_after();
      float* st = stage_out + q * (32 * 36);            // 32 rows x 32 cols, row stride 36 floats (16 B aligned)
      const int row_base = mb * BM + q * 32;
      const bool vec_c = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((tl.c_col & 3) == 0) &&
                         (nb * BN + BN <= p.N) && (row_base + 32 <= p.M);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        // EPI: the mask rows (or, without a mask, the old C values of a beta != 0 product) of this 32 x 32 block are
        // requested before the TMEM read and the staging pass, so their global-memory latency is paid once per block
        // instead of once per row; one register array serves both uses
        float4 pre[EPI ? 8 : 1];
        bool pre_old = false, pre_mask = false;
        if constexpr (EPI) {
          if (vec_c && !use_atomics) {
            const int c4p = (lane & 7) * 4, rsubp = lane >> 3;
            const int gcolp = nb * BN + c0 + c4p;
            pre_mask = ep.mask && ((ep.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.mask) & 15) == 0);
            pre_old = !pre_mask && p.beta != 0.f;
            const float* src = pre_mask ? ep.mask + (int64_t)(row_base + rsubp) * ep.ldmask + gcolp
                                        : p.C + (int64_t)(row_base + rsubp) * p.ldc + tl.c_col + gcolp;
            const int64_t ld = pre_mask ? ep.ldmask : p.ldc;
            if (pre_mask || pre_old) {
#pragma unroll
              for (int i = 0; i < 8; ++i) pre[i] = *reinterpret_cast<const float4*>(src + (int64_t)(4 * i) * ld);
            }
          }
        }
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
        tmem_ld32(taddr, r);
        tmem_ld_wait();
        // thread = row, r[j] = column j  ->  staging tile  ->  8 lanes per row, float4 each: coalesced 128 B rows
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(st + lane * 36 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                          __uint_as_float(r[j + 3]));
        __syncwarp();
        const int c4 = (lane & 7) * 4, rsub = lane >> 3;
        if constexpr (!EPI) {
          if (vec_c) {
            float* cbase = p.C + (int64_t)(row_base + rsub) * p.ldc + tl.c_col + nb * BN + c0 + c4;
            // all eight staged rows are read back before the first global store is issued (one shared-memory
            // round trip per block instead of eight dependent ones)
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[i] = *reinterpret_cast<const float4*>(st + (rsub + 4 * i) * 36 + c4);
              v[i].x *= p.alpha; v[i].y *= p.alpha; v[i].z *= p.alpha; v[i].w *= p.alpha;
            }
            if (use_atomics) {
#pragma unroll
              for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<float4*>(cbase + (int64_t)(4 * i) * p.ldc), v[i]);
            } else if (p.beta != 0.f) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float* c = cbase + (int64_t)(4 * i) * p.ldc;
                const float4 o = *reinterpret_cast<const float4*>(c);
                v[i].x = fmaf(p.beta, o.x, v[i].x); v[i].y = fmaf(p.beta, o.y, v[i].y);
                v[i].z = fmaf(p.beta, o.z, v[i].z); v[i].w = fmaf(p.beta, o.w, v[i].w);
                *reinterpret_cast<float4*>(c) = v[i];
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(cbase + (int64_t)(4 * i) * p.ldc) = v[i];
            }
          } else {
            for (int i = 0; i < 8; ++i) {
              const int row = row_base + rsub + 4 * i;
              if (row >= p.M) break;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int gc = nb * BN + c0 + c4 + e;
                if (gc >= p.N) break;
                float* c = p.C + (int64_t)row * p.ldc + tl.c_col + gc;
                const float v = p.alpha * st[(rsub + 4 * i) * 36 + c4 + e];
                if (use_atomics) {
                  atomicAdd(c, v);
                } else if (p.beta == 0.f) {
                  *c = v;
                } else {
                  *c = fmaf(p.beta, *c, v);
                }
              }
            }
          }
        } else {
          if (vec_c) {
            const int gcol = nb * BN + c0 + c4;
            float* cbase = p.C + (int64_t)(row_base + rsub) * p.ldc + tl.c_col + gcol;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI && ep.bias) {
              const float* bp = ep.bias + gcol;
              if ((reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) bv = __ldg(reinterpret_cast<const float4*>(bp));
              else bv = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
            }
            const bool vec_m = EPI && ep.mask && ((ep.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.mask) & 15) == 0);
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[i] = *reinterpret_cast<const float4*>(st + (rsub + 4 * i) * 36 + c4);
              v[i].x *= p.alpha; v[i].y *= p.alpha; v[i].z *= p.alpha; v[i].w *= p.alpha;
            }
            if (use_atomics) {
#pragma unroll
              for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<float4*>(cbase + (int64_t)(4 * i) * p.ldc), v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float* c = cbase + (int64_t)(4 * i) * p.ldc;
                float4 w = v[i];
                if (p.beta != 0.f) {
                  const float4 o = pre_old ? pre[i] : *reinterpret_cast<const float4*>(c);
                  w.x = fmaf(p.beta, o.x, w.x); w.y = fmaf(p.beta, o.y, w.y);
                  w.z = fmaf(p.beta, o.z, w.z); w.w = fmaf(p.beta, o.w, w.w);
                }
                w.x += bv.x; w.y += bv.y; w.z += bv.z; w.w += bv.w;
                if (ep.relu) {
                  w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f);
                }
                if (ep.mask) {
                  float4 m;
                  if (pre_mask) {
                    m = pre[i];
                  } else {
                    const float* mp = ep.mask + (int64_t)(row_base + rsub + 4 * i) * ep.ldmask + gcol;
                    if (vec_m) m = __ldg(reinterpret_cast<const float4*>(mp));
                    else m = make_float4(__ldg(mp), __ldg(mp + 1), __ldg(mp + 2), __ldg(mp + 3));
                  }
                  w.x = m.x > 0.f ? w.x : 0.f; w.y = m.y > 0.f ? w.y : 0.f;
                  w.z = m.z > 0.f ? w.z : 0.f; w.w = m.w > 0.f ? w.w : 0.f;
                }
                *reinterpret_cast<float4*>(c) = w;
              }
            }
          } else {
            for (int i = 0; i < 8; ++i) {
              const int row = row_base + rsub + 4 * i;
              if (row >= p.M) break;
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int gc = nb * BN + c0 + c4 + e;
                if (gc >= p.N) break;
                float* c = p.C + (int64_t)row * p.ldc + tl.c_col + gc;
                float v = p.alpha * st[(rsub + 4 * i) * 36 + c4 + e];
                if (use_atomics) {
                  atomicAdd(c, v);
                } else {
                  if (p.beta != 0.f) v = fmaf(p.beta, *c, v);
                  if (EPI && ep.bias) v += __ldg(ep.bias + gc);
                  if (EPI && ep.relu) v = fmaxf(v, 0.f);
                  if (EPI && ep.mask) v = __ldg(ep.mask + (int64_t)row * ep.ldmask + gc) > 0.f ? v : 0.f;
                  *c = v;
                }
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

__global__ void scale_matrix_tc_kernel(int M, int N, float* C, int64_t ldc, float beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  float* c = C + (i / N) * ldc + (i % N);
  *c = (beta == 0.f) ? 0.f : *c * beta;
}

// Tuning knobs for small products (read once): GS_TC_MIN_KB = fewest 64-wide k-blocks a K split may be left with
// (default 4), GS_TC_SHRINK_BN = 1 lets small problems (K <= 2048) use narrower accumulator tiles so that more CTAs
// share them.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int tune_min_kb() {
  static int v = -1;
  if (v < 0) v = std::max(1, env_int("GS_TC_MIN_KB", 4));
  return v;
}
static int tune_shrink_bn() {
  static int v = -1;
  if (v < 0) v = env_int("GS_TC_SHRINK_BN", 1);
  return v;
}

template <int BN, int NPASS>
static int launch(Params& p, const Epi& ep, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  constexpr bool kWithLo = NPASS == 3;
  constexpr size_t smem = (size_t)kStages * kPlanes * (BM * BK * 2 + BN * BK * 2) + 4 * 32 * 36 * 4 + 1024;
  {
    // B tile image into the caller's workspace
    const int kblocks_total = (p.K + BK - 1) / BK;
    const int tiles_n = (p.N + BN - 1) / BN;
    const size_t need = (size_t)tiles_n * kblocks_total * kPlanes * BN * BK * 2;
    if (!workspace || (size_t)workspace_bytes < need) {
      set_error_msg("gs_gemm_f32: workspace too small for the BF16 tile image of B (see gs_gemm_workspace_bytes)");
      return GS_ENOSPC;
    }
    constexpr size_t pack_smem = (size_t)2 * BN * BK * 2;     // lo plane address is formed even when unused
    static bool pack_configured = false;
    if (!pack_configured && pack_smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(pack_b_kernel<BN, kWithLo>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)pack_smem);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(pack_b)", e);
        return (int)e;
      }
      pack_configured = true;
    }
    pack_b_kernel<BN, kWithLo><<<tiles_n * kblocks_total, kProducerThreads, pack_smem, st>>>(
        p.B, p.ldb, p.tb, p.N, p.K, kblocks_total, reinterpret_cast<uint8_t*>(workspace));
    const int rc = finish_launch("pack_b");
    if (rc) return rc;
    p.Bimg = reinterpret_cast<const uint8_t*>(workspace);
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, NPASS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tc_kernel<BN, NPASS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tc)", e);
      return (int)e;
    }
    configured = true;
  }
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + BN - 1) / BN;
  const int64_t mn_tiles = (int64_t)p.tiles_m * p.tiles_n;
  const int kblocks = (p.K + BK - 1) / BK;
  int64_t total;
  if (p.seg) {
    // grouped: C is pre-zeroed by the caller and accumulated with atomics; split every group along K until the
    // work list covers the machine about twice (groups have similar K: class batches of <= 256 targets)
    const int64_t base = mn_tiles * p.groups;
    int splits = (int)((2 * kNumSMs + base - 1) / base);
    const int avg_kb = std::max(1, kblocks / std::max(1, p.groups));
    if (splits > avg_kb / 2) splits = avg_kb / 2;
    if (splits < 1) splits = 1;
    p.splits = splits;
    p.kchunk = 0;
    total = base * splits;
  } else {
    // split K when the output tiles alone cannot occupy the SMs
    int splits = 1;
    const int min_kb = tune_min_kb();
    if (mn_tiles < kNumSMs && kblocks >= 2 * min_kb && !ep.bias && !ep.relu && !ep.mask) {
      splits = (int)((kNumSMs + mn_tiles - 1) / mn_tiles);
      if (splits > kblocks / min_kb) splits = kblocks / min_kb;
      if (splits < 1) splits = 1;
    }
    int kchunk = ((kblocks + splits - 1) / splits) * BK;
    splits = (p.K + kchunk - 1) / kchunk;
    p.splits = splits;
    p.kchunk = kchunk;
    if (splits > 1) {
      const int64_t n = (int64_t)p.M * p.N;
      scale_matrix_tc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.M, p.N, p.C, p.ldc, p.beta);
      const int rc = finish_launch("scale_matrix_tc");
      if (rc) return rc;
    }
    total = mn_tiles * splits;
  }
  const int grid = (int)(total < kNumSMs ? total : kNumSMs);
  const bool plain_store = p.splits == 1 && !p.seg;
  if (ep.bias || ep.relu || ep.mask || (plain_store && p.beta != 0.f))
    gemm_tc_kernel<BN, NPASS, true><<<grid, kThreads, smem, st>>>(p, ep);
  else gemm_tc_kernel<BN, NPASS, false><<<grid, kThreads, smem, st>>>(p, ep);
  return finish_launch("gemm_tc");
}

}  // namespace tc

// Called by gs_gemm_f32 when precision != 0.  Returns GS_ENOSYS for shapes better served by the SIMT path
// (tiny products where a 128-row tile would be mostly padding).
static inline int tc_bn_wide(int N) { return N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32)); }
// accumulator width: the widest tile that does not waste more than half of its columns -- narrowed for small problems
// until the output tiles cover about half of the SMs (a 909 x 256 product is 8 tiles at BN = 256 but 64 at BN = 32;
// the A conversion repeated per column tile is negligible at that size, the serial epilogue per CTA is not)
static inline int tc_bn(int M, int N, int K) {
  int bn = tc_bn_wide(N);
  // only for products whose A operand is small: every extra column tile converts A again, and a long K is better
  // covered by the K split (the N'^2-deep dW product of PGE must stay at one column tile)
  if (!tc::tune_shrink_bn() || K > 2048) return bn;
  const int64_t tm = (M + tc::BM - 1) / tc::BM;
  while (bn > 32 && tm * ((N + bn - 1) / bn) < kNumSMs / 2 && (N + bn / 2 - 1) / (bn / 2) > (N + bn - 1) / bn) bn >>= 1;
  return bn;
}

bool gemm_tc_covers(int M, int N, int K) {
  return !(K < 32 || N < 16 || (int64_t)M * N * K < (int64_t)1 << 22);
}

int64_t gemm_tc_workspace_bytes(int M, int N, int K, int precision) {
  if (precision == 0 || !gemm_tc_covers(M, N, K)) return 0;
  const int bn = tc_bn_wide(N);      // upper bound over the tile widths tc_bn(M, N) may pick (same image size for all)
  const int64_t planes = precision == 1 ? 2 : 1;
  return (int64_t)((N + bn - 1) / bn) * ((K + tc::BK - 1) / tc::BK) * planes * bn * tc::BK * 2;
}

int gemm_tc_dispatch(int ta, int tb, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float beta, float* C, int64_t ldc, int precision, void* workspace,
                     int64_t workspace_bytes, const float* bias, int relu, const float* mask, int64_t ldmask,
                     cudaStream_t st) {
  if (!gemm_tc_covers(M, N, K)) return GS_ENOSYS;
  tc::Params p{ta, tb, M, N, K, alpha, beta, A, lda, B, ldb, nullptr, C, ldc, 1, K, 0, 0, nullptr, nullptr, 0};
  const tc::Epi ep{bias, relu, mask, ldmask};
  const bool three = precision == 1;
  switch (tc_bn(M, N, K)) {
    case 256:
      return three ? tc::launch<256, 3>(p, ep, workspace, workspace_bytes, st) : tc::launch<256, 1>(p, ep, workspace, workspace_bytes, st);
    case 128:
      return three ? tc::launch<128, 3>(p, ep, workspace, workspace_bytes, st) : tc::launch<128, 1>(p, ep, workspace, workspace_bytes, st);
    case 64:
      return three ? tc::launch<64, 3>(p, ep, workspace, workspace_bytes, st) : tc::launch<64, 1>(p, ep, workspace, workspace_bytes, st);
    default:
      return three ? tc::launch<32, 3>(p, ep, workspace, workspace_bytes, st) : tc::launch<32, 1>(p, ep, workspace, workspace_bytes, st);
  }
}


// Grouped K-segmented product on the tensor cores: C[:, out_block[g]*N ...] += A[seg[g]:seg[g+1]]^T B[seg[g]:seg[g+1]].
// Requirements (checked by the caller): every seg[g] is a multiple of 64 (gs_sampler_set_align(64)), C zeroed.
int gemm_tc_grouped_dispatch(int G, const int32_t* seg, const int32_t* out_block, int M, int N, int K_total,
                             const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                             int precision, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (!gemm_tc_covers(M, N, K_total) || M < 8) return GS_ENOSYS;
  tc::Params p{1, 0, M, N, K_total, 1.f, 0.f, A, lda, B, ldb, nullptr, C, ldc, 1, K_total, 0, 0, seg, out_block, G};
  const tc::Epi ep{nullptr, 0, nullptr, 0};
  const bool three = precision == 1;
  switch (tc_bn_wide(N)) {
    case 256:
      return three ? tc::launch<256, 3>(p, ep, workspace, workspace_bytes, st) : tc::launch<256, 1>(p, ep, workspace, workspace_bytes, st);
    case 128:
      return three ? tc::launch<128, 3>(p, ep, workspace, workspace_bytes, st) : tc::launch<128, 1>(p, ep, workspace, workspace_bytes, st);
    case 64:
      return three ? tc::launch<64, 3>(p, ep, workspace, workspace_bytes, st) : tc::launch<64, 1>(p, ep, workspace, workspace_bytes, st);
    default:
      return three ? tc::launch<32, 3>(p, ep, workspace, workspace_bytes, st) : tc::launch<32, 1>(p, ep, workspace, workspace_bytes, st);
  }
}

}  // namespace gs
