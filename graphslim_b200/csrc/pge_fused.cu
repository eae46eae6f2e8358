// Fused PGE layer-2 pipeline on tcgen05 + TMEM + TMA (sm_100a): the three N'^2-deep products of the pairwise adjacency
// MLP (graphslim/models/parametrized_adj.py:57-71 and their autograd) without materialising any N'^2 x h operand
// except the one activation the backward needs (Y2).
//
//   forward   Y2 = relu(bn1(Pa[j] + Pb[i])) W2^T            H1 is GENERATED in the A-producer (never in HBM); the
//                                                            epilogue writes Y2 and accumulates the BatchNorm-2 column
//                                                            sums (sum y, sum y^2) on the way out
//   backward  dH1 = dY2 W2, masked by H1 > 0, reduced        dY2 = bn2'(relu'(.) dE w3) is COMPUTED in the A-producer from
//             to Ga[j] = sum_i, Gb[i] = sum_j                the raw Y2 tile that TMA (cp.async.bulk.tensor) drops into the
//                                                            operand stage, converted in place; the epilogue re-generates
//                                                            the ReLU mask and reduces the tile, dH1 never reaches HBM
//   backward  dW2 = dY2^T H1   (K = N'^2)                    both operands produced on chip (dY2 as above, H1 generated),
//                                                            used as MN-major UMMA operands; the whole h x h result stays
//                                                            in TMEM for the CTA's share of the pair rows
//
// HBM traffic per outer step at the arxiv shape (N' = 909, h = 256; one N'^2 x h fp32 array = 846 MB):
//   forward 1 write (Y2) + 1 read (layer 3), backward 3 reads of Y2  ->  4.2 GB  (was 13.5 GB: H1, dY2, dH1 written and
//   re-read, Y2 read four times).
//
// Pair rows are tiled as BI x BJ = 16 x 8 blocks of (i, j) (tile row u = il*8 + jl), not as 128 consecutive rows: a
// tile then needs only 8 rows of Pa and 16 rows of Pb to generate H1, a raw Y2 tile is one 3-D TMA box per 32 columns
// (cols, j, i), and the backward epilogue can reduce the tile over il (-> Ga, 16x fewer atomics) and over jl (-> Gb,
// kept in registers across the consecutive tiles of one i-block).  Every CTA owns a contiguous range of tiles.
//
// BF16 multiplicands with the hi/lo split of gemm_tc.cu (precision 1: hi*hi + hi*lo + lo*hi; precision 2: hi*hi),
// fp32 accumulation in TMEM.  Warp roles (14 warps, one persistent CTA per SM): 0-7 producers, 8-11 epilogue
// (TMEM lane quarter = warp % 4), 12 MMA issuer (one lane), 13 TMA issuer (one lane).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gs {
namespace pf {

using namespace gs::ptx;

constexpr int BI = 16, BJ = 8;       // tile = BI x BJ pair rows (i, j), row u = il * BJ + jl
constexpr int BM = BI * BJ;          // 128 = UMMA M
constexpr int BK = 64;               // 64 bf16 = one 128-byte swizzle row
constexpr int kProducerWarps = 8, kProducerThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kMmaWarp = 12, kTmaWarp = 13;
constexpr int kThreads = 14 * 32;
constexpr uint32_t kPlaneA = BM * BK * 2;          // 16 KB: one bf16 plane of the A stage == one raw 32-column box
constexpr uint32_t kStageA = 2 * kPlaneA;          // hi | lo  ==  raw cols 0-31 | raw cols 32-63 (in-place conversion)
constexpr uint32_t kStagingBytes = 4 * 32 * 36 * 4;   // epilogue transposition tiles, one per warp
constexpr uint32_t kGaBytes = BJ * 256 * 4;           // the tile's Ga block (BJ x H floats), shared by the epilogue warps
constexpr int DW_BI = 4;                             // dW kernel: a k-stage is DW_BI x BJ = 32 pair rows
constexpr int DW_ROWS = DW_BI * BJ;
constexpr int kDwStages = 3;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// contiguous share of `num` work items for this CTA
__device__ __forceinline__ void cta_range(int num, int& t0, int& t1) {
  t0 = (int)(((int64_t)num * blockIdx.x) / gridDim.x);
  t1 = (int)(((int64_t)num * (blockIdx.x + 1)) / gridDim.x);
}

struct Geom {
  int n, n_i, i_first;   // pair rows (i, j): i in [i_first, i_first + n_i) (a rank's slice), j in [0, n)
  int tiles_j;           // ceil(n / BJ)
  int num_tiles;
};

struct Bn {              // BatchNorm affine pieces of one layer (h values each)
  const float *mean, *rstd, *gamma, *beta;
};

// BatchNorm-1 + ReLU of a generated layer-1 pre-activation with the affine pieces folded per column:
//   relu(gamma * ((a + p - mean) * rstd) + beta) = relu((a + p) * grs + off),  grs = gamma * rstd, off = beta - mean * grs.
// Every generator of H1 (forward producer, dW producer, the mask of the dX epilogue) uses this one form.
struct H1Consts {
  float4 grs, off;
};
__device__ __forceinline__ H1Consts load_h1_consts(const Bn& bn1, int c) {
  const float4 mu = ld4(bn1.mean + c), rs = ld4(bn1.rstd + c), g = ld4(bn1.gamma + c), b = ld4(bn1.beta + c);
  H1Consts k;
  k.grs = make_float4(g.x * rs.x, g.y * rs.y, g.z * rs.z, g.w * rs.w);
  k.off = make_float4(fmaf(-mu.x, k.grs.x, b.x), fmaf(-mu.y, k.grs.y, b.y), fmaf(-mu.z, k.grs.z, b.z),
                      fmaf(-mu.w, k.grs.w, b.w));
  return k;
}
__device__ __forceinline__ float h1_pre(float a, float p, float grs, float off) { return fmaf(a + p, grs, off); }
__device__ __forceinline__ float4 h1_value4(const float4& a, const float4& p, const H1Consts& k) {
  return make_float4(fmaxf(h1_pre(a.x, p.x, k.grs.x, k.off.x), 0.f), fmaxf(h1_pre(a.y, p.y, k.grs.y, k.off.y), 0.f),
                     fmaxf(h1_pre(a.z, p.z, k.grs.z, k.off.z), 0.f), fmaxf(h1_pre(a.w, p.w, k.grs.w, k.off.w), 0.f));
}

// d loss / d Y2 of one element = BatchNorm-2 backward of (relu'(bn2(y)) dE w3)   (pge_bn2_bwd_apply_kernel), folded:
//   xh = y * rs + nm (nm = -mean * rs);  on = gamma * xh + beta > 0;
//   dy2 = c0 * (on * dE * w3 - s1/M - xh * s2/M) = fma(xh, na2, on ? fma(dE, wc, na1) : na1)
// with c0 = gamma * rs, wc = w3 * c0, na1 = -(s1/M) c0, na2 = -(s2/M) c0.
struct Dy2Consts {
  float4 rs, nm, g, b, wc, na1, na2;
};
__device__ __forceinline__ Dy2Consts load_dy2_consts(const Bn& bn2, const float* w3, const float* s1, const float* s2,
                                                     float inv_count, int c) {
  const float4 mu = ld4(bn2.mean + c), w = ld4(w3 + c), a1 = ld4(s1 + c), a2 = ld4(s2 + c);
  Dy2Consts k;
  k.rs = ld4(bn2.rstd + c); k.g = ld4(bn2.gamma + c); k.b = ld4(bn2.beta + c);
  k.nm = make_float4(-mu.x * k.rs.x, -mu.y * k.rs.y, -mu.z * k.rs.z, -mu.w * k.rs.w);
  const float4 c0 = make_float4(k.g.x * k.rs.x, k.g.y * k.rs.y, k.g.z * k.rs.z, k.g.w * k.rs.w);
  k.wc = make_float4(w.x * c0.x, w.y * c0.y, w.z * c0.z, w.w * c0.w);
  k.na1 = make_float4(-a1.x * inv_count * c0.x, -a1.y * inv_count * c0.y, -a1.z * inv_count * c0.z, -a1.w * inv_count * c0.w);
  k.na2 = make_float4(-a2.x * inv_count * c0.x, -a2.y * inv_count * c0.y, -a2.z * inv_count * c0.z, -a2.w * inv_count * c0.w);
  return k;
}
__device__ __forceinline__ float dy2_value(float y, float de, float rs, float nm, float g, float b, float wc, float na1,
                                           float na2) {
  const float xh = fmaf(y, rs, nm);
  const float sel = (fmaf(g, xh, b) > 0.f) ? fmaf(de, wc, na1) : na1;
  return fmaf(xh, na2, sel);
}
__device__ __forceinline__ float4 dy2_value4(const float4& y, float de, const Dy2Consts& k) {
  return make_float4(dy2_value(y.x, de, k.rs.x, k.nm.x, k.g.x, k.b.x, k.wc.x, k.na1.x, k.na2.x),
                     dy2_value(y.y, de, k.rs.y, k.nm.y, k.g.y, k.b.y, k.wc.y, k.na1.y, k.na2.y),
                     dy2_value(y.z, de, k.rs.z, k.nm.z, k.g.z, k.b.z, k.wc.z, k.na1.z, k.na2.z),
                     dy2_value(y.w, de, k.rs.w, k.nm.w, k.g.w, k.b.w, k.wc.w, k.na1.w, k.na2.w));
}

// four fp32 values -> 8 bytes of the hi plane (+ 8 bytes of the lo plane) at a swizzled operand offset
template <bool kWithLo>
__device__ __forceinline__ void store_split4(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t off, const float4& v) {
  uint2 ph, pl;
  split_bf16x2(v.x, v.y, ph.x, pl.x);
  split_bf16x2(v.z, v.w, ph.y, pl.y);
  *reinterpret_cast<uint2*>(hi_plane + off) = ph;
  if (kWithLo) *reinterpret_cast<uint2*>(lo_plane + off) = pl;
}

// ------------------------------------------------------------------------------------------ W2 tile image
// BF16 hi/lo image of the reused operand in the K-major SWIZZLE_128B stage layout: [k block][plane][h rows x 128 B];
// operand row nr, depth k is W[nr*ldw + k] (trans == 0: the forward's W2, rows = outputs) or W[k*ldw + nr] (trans == 1:
// the dH1 product, rows = inputs).
__global__ void pack_w_image_kernel(const float* __restrict__ W, int64_t ldw, int h, int trans, int planes,
                                    uint8_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= h * h) return;
  const int nr = idx / h, k = idx % h;
  const float x = trans ? W[(int64_t)k * ldw + nr] : W[(int64_t)nr * ldw + k];
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  const int kb = k / BK, kk = k % BK;
  const size_t plane_bytes = (size_t)h * BK * 2;
  uint8_t* base = img + (size_t)kb * planes * plane_bytes;
  *reinterpret_cast<__nv_bfloat16*>(base + sw128_offset(nr, kk)) = hi;
  if (planes == 2) *reinterpret_cast<__nv_bfloat16*>(base + plane_bytes + sw128_offset(nr, kk)) = lo;
}

// ------------------------------------------------------------------------------------------ shared pieces
template <int H, int NPASS>
struct Cfg {
  static constexpr bool kWithLo = NPASS == 3;
  static constexpr int kPlanes = kWithLo ? 2 : 1;
  static constexpr int KB = H / BK;                  // k-blocks of a tile
  static constexpr int NB = H / 32;                  // 32-column blocks of an accumulator
  static constexpr uint32_t kPlaneB = H * BK * 2;    // one bf16 plane of the B stage (H operand rows)
  static constexpr uint32_t kStageBytes = kStageA + kPlanes * kPlaneB;
  static constexpr int kStages = (H == 256) ? 2 : 3;
  static constexpr uint32_t kTmemCols = 2 * H;       // double-buffered 128 x H accumulator
  static constexpr uint32_t kIdesc = make_idesc_bf16(BM, H, 0, 0);
  static_assert(H == 128 || H == 256, "PGE hidden width must be 128 or 256");
};

struct PipeBars {
  uint64_t raw[3], full[3], empty[3], tfull[2], tempty[2];
  uint32_t tmem_holder;
};

// MMA issuer of the two row-tile kernels: tile by tile, k-block by k-block, NPASS x 4 UMMAs (128 x H x 16) per stage
template <int H, int NPASS>
__device__ __forceinline__ void mma_role_rowtiles(uint8_t* smem, PipeBars& bars, uint32_t tmem_base, int t0, int t1,
                                                  int lane) {
  using C = Cfg<H, NPASS>;
  uint32_t it = 0, tcount = 0;
  for (int t = t0; t < t1; ++t, ++tcount) {
    const int acc = tcount & 1;
    const uint32_t acc_ph = (tcount >> 1) & 1;
    mbar_wait(smem_u32(&bars.tempty[acc]), acc_ph ^ 1);
    tc_fence_after();
    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * H);
    for (int kb = 0; kb < C::KB; ++kb, ++it) {
      const int s = it % C::kStages;
      const uint32_t ph = (it / C::kStages) & 1;
      mbar_wait(smem_u32(&bars.full[s]), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa_hi = smem_u32(smem + s * C::kStageBytes);
        const uint32_t sa_lo = sa_hi + kPlaneA;
        const uint32_t sb_hi = sa_hi + kStageA;
        const uint32_t sb_lo = sb_hi + C::kPlaneB;
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          const uint32_t a_base = (pass == 2) ? sa_lo : sa_hi;
          const uint32_t b_base = (pass == 1) ? sb_lo : sb_hi;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            umma_bf16(tmem_d, make_desc_k_sw128(a_base + kk * 32), make_desc_k_sw128(b_base + kk * 32), C::kIdesc,
                      (kb | pass | kk) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&bars.empty[s]));
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(smem_u32(&bars.tfull[acc]));
    __syncwarp();
  }
}

template <int STAGES>
__device__ __forceinline__ void init_bars(PipeBars& bars, uint32_t full_count, uint32_t tempty_count) {
  for (int s = 0; s < STAGES; ++s) {
    mbar_init(smem_u32(&bars.raw[s]), 1);
    mbar_init(smem_u32(&bars.full[s]), full_count);
    mbar_init(smem_u32(&bars.empty[s]), 1);
  }
  for (int a = 0; a < 2; ++a) {
    mbar_init(smem_u32(&bars.tfull[a]), 1);
    mbar_init(smem_u32(&bars.tempty[a]), tempty_count);
  }
  fence_barrier_init();
}

// epilogue helper: TMEM block (this warp's 32 lanes x 32 columns) -> the warp's staging tile (row stride 36 floats)
__device__ __forceinline__ void stage_block(float* st, int lane, const uint32_t (&r)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    *reinterpret_cast<float4*>(st + lane * 36 + j) =
        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
}

// =============================================================================================================
// forward: Y2 = relu(bn1(Pa[j] + Pb[i])) W2^T, + column sums of Y2
// =============================================================================================================
struct FwdParams {
  Geom g;
  const float *Pa, *Pb;      // n x H
  Bn bn1;
  const uint8_t* Bimg;       // image of W2 (operand rows = outputs)
  float* Y2;                 // (n_i * n) x H, row (li, j) at (li*n + j)*H
  double* stats;             // [sum y (H) | sum y^2 (H)], accumulated with atomics
};

// Warp roles of the forward (18 warps): 8 producer warps, 8 epilogue warps (two per TMEM lane quarter, alternating
// 32-column blocks), the MMA issuer and the TMA issuer.  ncu showed the 4-warp epilogue (v1) and then the 4-warp producer
// (v2: exposed L2 latency of the Pb rows) starving the tensor pipe in turn; both roles get 8 warps.
constexpr int kFwdProducerWarps = 8;
constexpr int kFwdEpiWarps = 8;
constexpr int kFwdMmaWarp = 16, kFwdTmaWarp = 17;
constexpr int kFwdThreads = 18 * 32;
constexpr uint32_t kFwdStagingBytes = kFwdEpiWarps * 32 * 32 * 4;     // dense swizzled 32 x 32 tiles, one per warp

template <int H, int NPASS, bool kTmaStore>
__global__ void __launch_bounds__(kFwdThreads, 1) pge_l2_fwd_kernel(const __grid_constant__ CUtensorMap map_y2, FwdParams p) {
  using C = Cfg<H, NPASS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  __shared__ __align__(8) PipeBars bars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t0, t1;
  cta_range(p.g.num_tiles, t0, t1);

  if (threadIdx.x == 0) init_bars<C::kStages>(bars, kFwdProducerWarps + 1, kFwdEpiWarps * 32);
  if (warp == kFwdMmaWarp) tmem_alloc(smem_u32(&bars.tmem_holder), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_holder;

  if (warp < kFwdProducerWarps) {
    // ============================== producers: generate the H1 tile ==============================
    // thread = (c4, jl, ilb): tile rows u = jl + 8 il for il = ilb + 2 i; one Pa vector and 8 Pb vectors per k-block, all
    // requested before the stage wait so their L2 latency overlaps it
    const int tid = threadIdx.x;
    const int c4 = tid & 15, jl = (tid >> 4) & 7, ilb = tid >> 7;
    uint32_t it = 0;
    for (int t = t0; t < t1; ++t) {
      const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
      const int j = jb * BJ + jl;
      const bool jv = j < p.g.n;
      const float* pa_row = p.Pa + (int64_t)(jv ? j : 0) * H;
      const float* pb_rows = p.Pb + (int64_t)(p.g.i_first + ib * BI) * H;
      const int ni = min(BI, p.g.n_i - ib * BI);                 // valid il of this tile (>= 1)
#pragma unroll 1
      for (int kb = 0; kb < C::KB; ++kb, ++it) {
        const int c = kb * BK + c4 * 4;
        const H1Consts k1 = load_h1_consts(p.bn1, c);
        const float4 pa = ld4(pa_row + c);
        float4 pb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int il = ilb + 2 * i;
          pb[i] = ld4(pb_rows + (int64_t)(il < ni ? il : 0) * H + c);
        }
        const int s = it % C::kStages;
        const uint32_t ph = (it / C::kStages) & 1;
        mbar_wait(smem_u32(&bars.empty[s]), ph ^ 1);
        uint8_t* sa_hi = smem + s * C::kStageBytes;
        uint8_t* sa_lo = sa_hi + kPlaneA;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int il = ilb + 2 * i;
          float4 o = h1_value4(pa, pb[i], k1);
          if (!(jv && il < ni)) o = make_float4(0.f, 0.f, 0.f, 0.f);
          store_split4<C::kWithLo>(sa_hi, sa_lo, sw128_offset(il * BJ + jl, c4 * 4), o);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars.full[s]));
      }
    }
  } else if (warp == kFwdTmaWarp) {
    // ============================== TMA: the W2 image block of every k-block ==============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = t0; t < t1; ++t) {
        for (int kb = 0; kb < C::KB; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(smem_u32(&bars.empty[s]), ph ^ 1);
          const uint32_t bar = smem_u32(&bars.full[s]);
          mbar_arrive_expect_tx(bar, C::kPlanes * C::kPlaneB);
          bulk_g2s(smem_u32(smem + s * C::kStageBytes + kStageA), p.Bimg + (size_t)kb * C::kPlanes * C::kPlaneB,
                   C::kPlanes * C::kPlaneB, bar);
        }
      }
    }
  } else if (warp == kFwdMmaWarp) {
    mma_role_rowtiles<H, NPASS>(smem, bars, tmem_base, t0, t1, lane);
  } else {
    // ============================== epilogue: store Y2, accumulate its column sums ==============================
    constexpr int NBW = C::NB / 2;                       // column blocks per warp
    const int ew = warp - kFwdProducerWarps;             // 0..7
    const int q = warp & 3, half = ew >> 2;              // TMEM lane quarter; parity of the blocks this warp handles
    float* st = staging + ew * (32 * 32);
    // fp32 partial sums per block, folded into doubles every 8 tiles (DADD is slow on this part)
    float ps1[NBW], ps2[NBW];
    double d1[NBW], d2[NBW];
#pragma unroll
    for (int b = 0; b < NBW; ++b) {
      ps1[b] = ps2[b] = 0.f;
      d1[b] = d2[b] = 0.0;
    }
    uint32_t tcount = 0;
    const int ch = lane & 7, rsub = lane >> 3;
    for (int t = t0; t < t1; ++t, ++tcount) {
      const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
      const int acc = tcount & 1;
      const uint32_t acc_ph = (tcount >> 1) & 1;
      mbar_wait(smem_u32(&bars.tfull[acc]), acc_ph);
      tc_fence_after();
#pragma unroll
      for (int bi = 0; bi < NBW; ++bi) {
        const int b = 2 * bi + half;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * H + b * 32), r);
        tmem_ld_wait();
        // dense 32 x 32 tile, 16-byte chunk j of row u stored at chunk j ^ (u & 7): conflict-free for the row writes
        // (thread = row), the column reads (lane = column) and the row reads of the coalesced store
#pragma unroll
        for (int jc = 0; jc < 8; ++jc)
          *reinterpret_cast<float4*>(st + lane * 32 + ((jc ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(r[4 * jc]), __uint_as_float(r[4 * jc + 1]), __uint_as_float(r[4 * jc + 2]),
                          __uint_as_float(r[4 * jc + 3]));
        if constexpr (kTmaStore) {
          // The tile IS the SWIZZLE_128B image of a (4 il x 8 jl x 32 columns) box of Y2 (rows of 128 bytes, 16-byte chunk
          // ^ row & 7, 1 KB-aligned): one TMA store per block instead of 8 LDS.128 + 8 STG.128 per lane.  ncu had the LSU
          // data pipe of this kernel at 90 % (11.9 k wavefronts per tile against 6.1 k MMA cycles); rows / columns outside
          // the slice are clipped by the tensor map.
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_y2, smem_u32(st), b * 32, jb * BJ, ib * BI + 4 * q);
            bulk_commit_group();
          }
        } else {
          __syncwarp();
        }
        // rows outside the slice were generated as zeros, so they add nothing to the sums
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const float x = st[u * 32 + ((((lane >> 2) ^ (u & 7)) << 2) | (lane & 3))];
          a1 += x;
          a2 = fmaf(x, x, a2);
        }
        ps1[bi] += a1;
        ps2[bi] += a2;
        if constexpr (kTmaStore) {
          if (lane == 0) bulk_wait_group_read0();           // the TMA engine has read the tile: it may be rewritten
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = rsub + 4 * k, u = q * 32 + row;
            const int li = ib * BI + (u >> 3), j = jb * BJ + (u & 7);
            if (li < p.g.n_i && j < p.g.n)
              *reinterpret_cast<float4*>(p.Y2 + ((int64_t)li * p.g.n + j) * H + b * 32 + ch * 4) =
                  *reinterpret_cast<const float4*>(st + row * 32 + ((ch ^ (row & 7)) << 2));
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars.tempty[acc]));
      if ((tcount & 7) == 7) {
#pragma unroll
        for (int b = 0; b < NBW; ++b) {
          d1[b] += (double)ps1[b];
          d2[b] += (double)ps2[b];
          ps1[b] = ps2[b] = 0.f;
        }
      }
    }
    if constexpr (kTmaStore) {
      if (lane == 0) bulk_wait_group0();                    // every store of this warp has landed
      __syncwarp();
    }
    // column sums: the four lane quarters of each block parity are combined through shared memory, one double atomic
    // per column, statistic and CTA
    double* red = reinterpret_cast<double*>(staging);     // [8 warps][NBW][2][32] doubles = 16 KB of the 32 KB tiles
    named_bar_sync(1, kFwdEpiWarps * 32);
#pragma unroll
    for (int b = 0; b < NBW; ++b) {
      red[((ew * NBW + b) * 2 + 0) * 32 + lane] = d1[b] + (double)ps1[b];
      red[((ew * NBW + b) * 2 + 1) * 32 + lane] = d2[b] + (double)ps2[b];
    }
    named_bar_sync(1, kFwdEpiWarps * 32);
    for (int idx = ew * 32 + lane; idx < 2 * H; idx += kFwdEpiWarps * 32) {
      const int stat = idx / H, col = idx % H;
      const int b = col >> 5, l = col & 31;
      double v = 0.0;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) v += red[((((b & 1) * 4 + qq) * NBW + (b >> 1)) * 2 + stat) * 32 + l];
      atomicAdd(p.stats + idx, v);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kFwdMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// =============================================================================================================
// forward on CTA PAIRS (cta_group::2): one UMMA of M = 256 covers the tiles of both SMs of a TPC; each CTA keeps ITS HALF
// of the W2 image (128 of the 256 operand rows, all k-blocks, hi + lo planes: 128 KB) resident in shared memory for the
// whole kernel.  Against the single-CTA kernel this removes the 64 KB W2 refill per k-block and halves the B bytes an
// UMMA reads from shared memory (the single-CTA kernels are bound by shared-memory bandwidth).
// =============================================================================================================
template <int H, int NPASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) pge_l2_fwd2_kernel(FwdParams p) {
  using C = Cfg<H, NPASS>;
  constexpr int kFwdProducerWarps = 4;                  // this variant keeps the 14-warp layout (4 + 8 + MMA + TMA)
  constexpr int kStages2 = 2;
  constexpr uint32_t kBHalfPlane = (H / 2) * BK * 2;                      // one plane of this CTA's half of a k-block
  constexpr uint32_t kBHalf = C::KB * C::kPlanes * kBHalfPlane;           // resident: 128 KB at H = 256, 3 passes
  constexpr uint32_t kIdesc2 = make_idesc_bf16(256, H, 0, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                                     // [kb][plane][H/2 rows x 128 B]
  uint8_t* sA = smem + kBHalf;                                            // kStages2 x (hi | lo) A tiles
  float* staging = reinterpret_cast<float*>(sA + kStages2 * kStageA);
  __shared__ __align__(8) PipeBars bars;
  __shared__ __align__(8) uint64_t bar_bload;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_st = (p.g.num_tiles + 1) >> 1;                            // super tiles: tiles (2 st, 2 st + 1)
  const int st0 = (int)(((int64_t)num_st * pair) / npairs), st1 = (int)(((int64_t)num_st * (pair + 1)) / npairs);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(smem_u32(&bars.full[s]), 2 * kFwdProducerWarps);          // leader's copy: producers of both CTAs
      mbar_init(smem_u32(&bars.empty[s]), 1);                             // multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bars.tfull[a]), 1);                             // multicast commit
      mbar_init(smem_u32(&bars.tempty[a]), 2 * kFwdEpiWarps * 32);        // leader's copy: epilogue threads of both CTAs
    }
    mbar_init(smem_u32(&bar_bload), 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc_2cta(smem_u32(&bars.tmem_holder), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kTmaWarp && lane == 0) {
    // this CTA's half of the W2 image: rows [rank * H/2, +H/2) of every (k-block, plane) slab
    const uint32_t bar = smem_u32(&bar_bload);
    mbar_arrive_expect_tx(bar, kBHalf);
    for (int kb = 0; kb < C::KB; ++kb)
      for (int pl = 0; pl < C::kPlanes; ++pl)
        bulk_g2s(smem_u32(sB + (kb * C::kPlanes + pl) * kBHalfPlane),
                 p.Bimg + (size_t)(kb * C::kPlanes + pl) * C::kPlaneB + (size_t)rank * kBHalfPlane, kBHalfPlane, bar);
  }
  mbar_wait(smem_u32(&bar_bload), 0);
  cluster_sync_all();                                   // both halves of B resident, both CTAs' barriers initialised
  const uint32_t tmem_base = bars.tmem_holder;

  if (warp < kFwdProducerWarps) {
    const int tid = threadIdx.x;
    const int c4 = tid & 15, jl = tid >> 4;
    uint32_t it = 0;
    for (int st = st0; st < st1; ++st) {
      const int t = 2 * st + (int)rank;                 // may be one past the last tile: generated as zeros
      const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
      const int j = jb * BJ + jl;
      const int ni = min(BI, p.g.n_i - ib * BI);        // <= 0 for the phantom tile
      const bool jv = j < p.g.n && ni > 0;
      const float* pa_row = p.Pa + (int64_t)(jv ? j : 0) * H;
      const float* pb_rows = p.Pb + (int64_t)(p.g.i_first + (ni > 0 ? ib * BI : 0)) * H;
#pragma unroll 1
      for (int kb = 0; kb < C::KB; ++kb, ++it) {
        const int c = kb * BK + c4 * 4;
        const H1Consts k1 = load_h1_consts(p.bn1, c);
        const float4 pa = ld4(pa_row + c);
        float4 pb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pb[i] = ld4(pb_rows + (int64_t)(i < ni ? i : 0) * H + c);
        const int s = it % kStages2;
        const uint32_t ph = (it / kStages2) & 1;
        mbar_wait_cluster(smem_u32(&bars.empty[s]), ph ^ 1);
        uint8_t* sa_hi = sA + s * kStageA;
        uint8_t* sa_lo = sa_hi + kPlaneA;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pb[i] = ld4(pb_rows + (int64_t)(8 + i < ni ? 8 + i : 0) * H + c);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int il = half * 8 + i;
            float4 o = h1_value4(pa, pb[i], k1);
            if (!(jv && il < ni)) o = make_float4(0.f, 0.f, 0.f, 0.f);
            store_split4<C::kWithLo>(sa_hi, sa_lo, sw128_offset(il * BJ + jl, c4 * 4), o);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&bars.full[s]), 0);      // the leader's barrier
      }
    }
  } else if (warp == kMmaWarp) {
    if (rank == 0) {
      uint32_t it = 0, tcount = 0;
      for (int st = st0; st < st1; ++st, ++tcount) {
        const int acc = tcount & 1;
        const uint32_t acc_ph = (tcount >> 1) & 1;
        mbar_wait_cluster(smem_u32(&bars.tempty[acc]), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * H);
        for (int kb = 0; kb < C::KB; ++kb, ++it) {
          const int s = it % kStages2;
          const uint32_t ph = (it / kStages2) & 1;
          mbar_wait_cluster(smem_u32(&bars.full[s]), ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa_hi = smem_u32(sA + s * kStageA);
            const uint32_t sa_lo = sa_hi + kPlaneA;
            const uint32_t sb_hi = smem_u32(sB + (kb * C::kPlanes) * kBHalfPlane);
            const uint32_t sb_lo = sb_hi + kBHalfPlane;
#pragma unroll
            for (int pass = 0; pass < NPASS; ++pass) {
              const uint32_t a_base = (pass == 2) ? sa_lo : sa_hi;
              const uint32_t b_base = (pass == 1) ? sb_lo : sb_hi;
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk) {
                umma_bf16_2cta(tmem_d, make_desc_k_sw128(a_base + kk * 32), make_desc_k_sw128(b_base + kk * 32), kIdesc2,
                               (kb | pass | kk) ? 1u : 0u);
              }
            }
            umma_commit_2cta(smem_u32(&bars.empty[s]));          // frees the stage in both CTAs
          }
          __syncwarp();
        }
        if (lane == 0) umma_commit_2cta(smem_u32(&bars.tfull[acc]));
        __syncwarp();
      }
    }
  } else if (warp == kTmaWarp) {
    // nothing to stream: B is resident
  } else {
    constexpr int NBW = C::NB / 2;
    const int ew = warp - kFwdProducerWarps;
    const int q = warp & 3, half = ew >> 2;
    float* st_ = staging + ew * (32 * 32);
    float ps1[NBW], ps2[NBW];
    double d1[NBW], d2[NBW];
#pragma unroll
    for (int b = 0; b < NBW; ++b) {
      ps1[b] = ps2[b] = 0.f;
      d1[b] = d2[b] = 0.0;
    }
    uint32_t tcount = 0;
    const int ch = lane & 7, rsub = lane >> 3;
    for (int st = st0; st < st1; ++st, ++tcount) {
      const int t = 2 * st + (int)rank;
      const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
      const int acc = tcount & 1;
      const uint32_t acc_ph = (tcount >> 1) & 1;
      mbar_wait_cluster(smem_u32(&bars.tfull[acc]), acc_ph);
      tc_fence_after();
#pragma unroll
      for (int bi = 0; bi < NBW; ++bi) {
        const int b = 2 * bi + half;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * H + b * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int jc = 0; jc < 8; ++jc)
          *reinterpret_cast<float4*>(st_ + lane * 32 + ((jc ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(r[4 * jc]), __uint_as_float(r[4 * jc + 1]), __uint_as_float(r[4 * jc + 2]),
                          __uint_as_float(r[4 * jc + 3]));
        __syncwarp();
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const float x = st_[u * 32 + ((((lane >> 2) ^ (u & 7)) << 2) | (lane & 3))];
          a1 += x;
          a2 = fmaf(x, x, a2);
        }
        ps1[bi] += a1;
        ps2[bi] += a2;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int row = rsub + 4 * k, u = q * 32 + row;
          const int li = ib * BI + (u >> 3), j = jb * BJ + (u & 7);
          if (li < p.g.n_i && j < p.g.n)
            *reinterpret_cast<float4*>(p.Y2 + ((int64_t)li * p.g.n + j) * H + b * 32 + ch * 4) =
                *reinterpret_cast<const float4*>(st_ + row * 32 + ((ch ^ (row & 7)) << 2));
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive_cluster(smem_u32(&bars.tempty[acc]), 0);
      if ((tcount & 7) == 7) {
#pragma unroll
        for (int b = 0; b < NBW; ++b) {
          d1[b] += (double)ps1[b];
          d2[b] += (double)ps2[b];
          ps1[b] = ps2[b] = 0.f;
        }
      }
    }
    double* red = reinterpret_cast<double*>(staging);
    named_bar_sync(1, kFwdEpiWarps * 32);
#pragma unroll
    for (int b = 0; b < NBW; ++b) {
      red[((ew * NBW + b) * 2 + 0) * 32 + lane] = d1[b] + (double)ps1[b];
      red[((ew * NBW + b) * 2 + 1) * 32 + lane] = d2[b] + (double)ps2[b];
    }
    named_bar_sync(1, kFwdEpiWarps * 32);
    for (int idx = ew * 32 + lane; idx < 2 * H; idx += kFwdEpiWarps * 32) {
      const int stat = idx / H, col = idx % H;
      const int b = col >> 5, l = col & 31;
      double v = 0.0;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) v += red[((((b & 1) * 4 + qq) * NBW + (b >> 1)) * 2 + stat) * 32 + l];
      atomicAdd(p.stats + idx, v);
    }
  }

  tc_fence_before();
  cluster_sync_all();                                   // the peer's barriers / TMEM stay valid until both are done
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, C::kTmemCols);
  }
}

// =============================================================================================================
// backward 1: dH1 = dY2 W2 (dY2 computed in the producer from the TMA-loaded raw Y2 tile), then either stored or
// masked with H1 > 0 and reduced to Ga[j] = sum_i, Gb[i] = sum_j in the epilogue
// =============================================================================================================
struct BwdParams {
  Geom g;
  const float *Pa, *Pb;      // n x H
  Bn bn1, bn2;
  const float* dE;           // n_i * n (this slice's rows of d loss / d E)
  const float *w3, *s1, *s2; // layer-3 weight and the global BatchNorm-2 backward sums (H each)
  float inv_count;           // 1 / (number of pair rows of the whole batch)
  const uint8_t* Bimg;       // image of W2 (operand rows = inputs), dX kernel only
  float *Ga, *Gb;            // n x H each: fused reduction targets (atomics)
  float* dH1;                // (n_i * n) x H when the tile is stored instead
  float* dW2;                // H x H (dW kernel; atomics)
};

template <int H, int NPASS, bool kFusedReduce>
__global__ void __launch_bounds__(kThreads, 1)
pge_l2_bwd_dx_kernel(const __grid_constant__ CUtensorMap map_y2, BwdParams p) {
  using C = Cfg<H, NPASS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  float* ga_sm = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + kStagingBytes);
  __shared__ __align__(8) PipeBars bars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t0, t1;
  cta_range(p.g.num_tiles, t0, t1);

  if (threadIdx.x == 0) init_bars<C::kStages>(bars, kProducerWarps + 1, kEpiThreads);
  if (warp == kMmaWarp) tmem_alloc(smem_u32(&bars.tmem_holder), C::kTmemCols);
  if (warp == kTmaWarp && lane == 0) tma_prefetch_desc(&map_y2);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_holder;

  if (warp < kProducerWarps) {
    // ============================== producers: raw Y2 tile -> dY2 operand, in place ==============================
    const int tid = threadIdx.x;
    const int c4 = tid & 15, jl = (tid >> 4) & 7, ilb = tid >> 7;
    uint32_t it = 0;
    for (int t = t0; t < t1; ++t) {
      const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
      const int j = jb * BJ + jl;
      const bool jv = j < p.g.n;
      float de[8];
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int li = ib * BI + ilb + 2 * i;
        const bool valid = jv && li < p.g.n_i;
        de[i] = valid ? __ldg(p.dE + (int64_t)li * p.g.n + j) : 0.f;
        vmask |= (valid ? 1u : 0u) << i;
      }
#pragma unroll 1
      for (int kb = 0; kb < C::KB; ++kb, ++it) {
        const int c = kb * BK + c4 * 4;
        const Dy2Consts k = load_dy2_consts(p.bn2, p.w3, p.s1, p.s2, p.inv_count, c);
        const int s = it % C::kStages;
        const uint32_t ph = (it / C::kStages) & 1;
        uint8_t* sa_hi = smem + s * C::kStageBytes;
        uint8_t* sa_lo = sa_hi + kPlaneA;
        // raw layout: 32-column box 0 in the hi-plane region, box 1 in the lo-plane region, row u at u * 128 B
        const uint8_t* raw = sa_hi + (c4 >> 3) * kPlaneA + (c4 & 7) * 16;
        mbar_wait(smem_u32(&bars.raw[s]), ph);
        float4 y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = *reinterpret_cast<const float4*>(raw + ((tid >> 4) + 16 * i) * 128);
        __syncwarp();            // a warp owns whole rows: every raw byte of them is in registers before any is overwritten
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u = (tid >> 4) + 16 * i;
          float4 o = dy2_value4(y[i], de[i], k);
          if (!((vmask >> i) & 1u)) o = make_float4(0.f, 0.f, 0.f, 0.f);
          store_split4<C::kWithLo>(sa_hi, sa_lo, sw128_offset(u, c4 * 4), o);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars.full[s]));
      }
    }
  } else if (warp == kTmaWarp) {
    // ============================== TMA: raw Y2 boxes + the W2 image block ==============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = t0; t < t1; ++t) {
        const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
        for (int kb = 0; kb < C::KB; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(smem_u32(&bars.empty[s]), ph ^ 1);
          const uint32_t dst = smem_u32(smem + s * C::kStageBytes);
          const uint32_t braw = smem_u32(&bars.raw[s]);
          mbar_arrive_expect_tx(braw, kStageA);
          tma_load_3d(dst, &map_y2, kb * BK, jb * BJ, ib * BI, braw);
          tma_load_3d(dst + kPlaneA, &map_y2, kb * BK + 32, jb * BJ, ib * BI, braw);
          const uint32_t bfull = smem_u32(&bars.full[s]);
          mbar_arrive_expect_tx(bfull, C::kPlanes * C::kPlaneB);
          bulk_g2s(dst + kStageA, p.Bimg + (size_t)kb * C::kPlanes * C::kPlaneB, C::kPlanes * C::kPlaneB, bfull);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    mma_role_rowtiles<H, NPASS>(smem, bars, tmem_base, t0, t1, lane);
  } else {
    // ============================== epilogue ==============================
    const int q = warp & 3;
    float* st = staging + q * (32 * 36);
    uint32_t tcount = 0;
    if constexpr (!kFusedReduce) {
      const int c4 = (lane & 7) * 4, rsub = lane >> 3;
      for (int t = t0; t < t1; ++t, ++tcount) {
        const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
        const int acc = tcount & 1;
        mbar_wait(smem_u32(&bars.tfull[acc]), (tcount >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int b = 0; b < C::NB; ++b) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * H + b * 32), r);
          tmem_ld_wait();
          stage_block(st, lane, r);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int u = q * 32 + rsub + 4 * k;
            const int li = ib * BI + (u >> 3), j = jb * BJ + (u & 7);
            if (li < p.g.n_i && j < p.g.n)
              *reinterpret_cast<float4*>(p.dH1 + ((int64_t)li * p.g.n + j) * H + b * 32 + c4) =
                  *reinterpret_cast<const float4*>(st + (rsub + 4 * k) * 36 + c4);
          }
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&bars.tempty[acc]));
      }
    } else {
      // The raw 32 x 32 accumulator block is staged (thread = tile row u = 32 q + lane, il = 4 q + u / 8, jl = u % 8);
      // then LANE c OWNS COLUMN c: it regenerates the ReLU mask of its column for the warp's 32 rows from 8 Pa values,
      // 4 Pb values and the column's BatchNorm constants (coalesced scalar loads), and sums the masked entries over
      // the 8 jl of each il (-> Gb, kept in registers while the i-block lasts) and over the warp's 4 il for each jl
      // (-> Ga: shared-memory atomics into an 8 x H tile shared by the four warps, flushed once per tile with float4
      // global atomics).  Rows outside the slice carry dY2 = 0, hence exact zeros here.
      float gb_acc[C::NB][4];
#pragma unroll
      for (int b = 0; b < C::NB; ++b)
#pragma unroll
        for (int k = 0; k < 4; ++k) gb_acc[b][k] = 0.f;
      int cur_ib = -1;
      auto flush_gb = [&](int ibf) {
#pragma unroll
        for (int b = 0; b < C::NB; ++b)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int li = ibf * BI + 4 * q + k;
            if (li < p.g.n_i) atomicAdd(p.Gb + (int64_t)(p.g.i_first + li) * H + b * 32 + lane, gb_acc[b][k]);
            gb_acc[b][k] = 0.f;
          }
      };
      const int et = q * 32 + lane;           // epilogue thread id 0..127
      for (int idx = et; idx < BJ * H; idx += kEpiThreads) ga_sm[idx] = 0.f;
      // BatchNorm constants of this lane's column of every 32-column block: tile independent, loaded once
      float grs_c[C::NB], off_c[C::NB];
#pragma unroll
      for (int b = 0; b < C::NB; ++b) {
        const int c = b * 32 + lane;
        grs_c[b] = __ldg(p.bn1.gamma + c) * __ldg(p.bn1.rstd + c);
        off_c[b] = fmaf(-__ldg(p.bn1.mean + c), grs_c[b], __ldg(p.bn1.beta + c));
      }
      named_bar_sync(1, kEpiThreads);
      for (int t = t0; t < t1; ++t, ++tcount) {
        const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
        if (ib != cur_ib) {
          if (cur_ib >= 0) flush_gb(cur_ib);
          cur_ib = ib;
        }
        const float* pa_base = p.Pa + (int64_t)(jb * BJ) * H + lane;
        const float* pb_base = p.Pb + (int64_t)(p.g.i_first + ib * BI + 4 * q) * H + lane;
        const int nj = min(BJ, p.g.n - jb * BJ), ni = min(4, p.g.n_i - (ib * BI + 4 * q));   // valid rows (ni may be <= 0)
        // the 8 Pa and 4 Pb values of a block are requested one block ahead (they were the exposed latency of this role:
        // ncu showed the fused-reduction kernel 180 us behind the variant that only stores dH1)
        float pa_n[BJ], pb_n[4];
        auto fetch = [&](int c) {
#pragma unroll
          for (int jj = 0; jj < BJ; ++jj) pa_n[jj] = __ldg(pa_base + (int64_t)(jj < nj ? jj : 0) * H + c);
#pragma unroll
          for (int k = 0; k < 4; ++k) pb_n[k] = (ni > 0) ? __ldg(pb_base + (int64_t)(k < ni ? k : 0) * H + c) : 0.f;
        };
        fetch(0);
        const int acc = tcount & 1;
        mbar_wait(smem_u32(&bars.tfull[acc]), (tcount >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int b = 0; b < C::NB; ++b) {
          const int c = b * 32;
          float pa[BJ], pb[4];
#pragma unroll
          for (int jj = 0; jj < BJ; ++jj) pa[jj] = pa_n[jj];
#pragma unroll
          for (int k = 0; k < 4; ++k) pb[k] = pb_n[k];
          const float grs = grs_c[b], off = off_c[b];
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * H + c), r);
          if (b + 1 < C::NB) fetch(c + 32);
          tmem_ld_wait();
          stage_block(st, lane, r);
          __syncwarp();
          float ga[BJ];
#pragma unroll
          for (int jj = 0; jj < BJ; ++jj) ga[jj] = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float gsum = 0.f;
#pragma unroll
            for (int jj = 0; jj < BJ; ++jj) {
              float x = st[(k * 8 + jj) * 36 + lane];
              x = (h1_pre(pa[jj], pb[k], grs, off) > 0.f) ? x : 0.f;
              gsum += x;
              ga[jj] += x;
            }
            gb_acc[b][k] += gsum;
          }
#pragma unroll
          for (int jj = 0; jj < BJ; ++jj) atomicAdd(ga_sm + jj * H + c + lane, ga[jj]);
          __syncwarp();                         // the staging tile is rewritten by the next block
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&bars.tempty[acc]));
        named_bar_sync(1, kEpiThreads);         // all four warps have added their share of this tile's Ga block
        for (int idx = et; idx < BJ * H / 4; idx += kEpiThreads) {
          const float4 v = reinterpret_cast<float4*>(ga_sm)[idx];
          reinterpret_cast<float4*>(ga_sm)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int jj = idx / (H / 4);
          if (jb * BJ + jj < p.g.n)
            atomicAdd(reinterpret_cast<float4*>(p.Ga + (int64_t)(jb * BJ + jj) * H) + (idx % (H / 4)), v);
        }
        named_bar_sync(1, kEpiThreads);         // zeroed before the next tile's adds
      }
      if (cur_ib >= 0) flush_gb(cur_ib);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// =============================================================================================================
// backward 2: dW2 = dY2^T H1 over this CTA's share of the pair rows; both operands MN-major (stored [pair row][column])
// =============================================================================================================
template <int H, int NPASS>
struct DwCfg {
  static constexpr bool kWithLo = NPASS == 3;
  static constexpr int CB = H / 64;                       // 64-column sub-tiles of an operand
  static constexpr uint32_t kSub = 2 * DW_ROWS * 128;     // one sub-tile: hi plane (32 rows x 128 B) | lo plane = 8 KB
  static constexpr uint32_t kOperand = CB * kSub;         // 32 KB at H = 256 == the raw fp32 stage (32 rows x H)
  static constexpr uint32_t kStageBytes = 2 * kOperand;   // dY2 | H1
  static constexpr int MH = H / 128;                      // 128-row halves of the h_out x h_in result
  static constexpr uint32_t kTmemCols = MH * H;           // 512 (H = 256) / 128 (H = 128)
  static constexpr uint32_t kIdesc = make_idesc_bf16(128, H, 1, 1);
  static constexpr int F4 = H / 4;                        // float4 per row
  static constexpr int kRowStep = 256 / F4;               // rows covered by one pass of the 256 producer threads
  static constexpr int kIters = DW_ROWS / kRowStep;       // 8 (H = 256) / 4 (H = 128)
};

template <int H, int NPASS>
__global__ void __launch_bounds__(kThreads, 1)
pge_l2_bwd_dw_kernel(const __grid_constant__ CUtensorMap map_y2, BwdParams p) {
  using C = DwCfg<H, NPASS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(smem + kDwStages * C::kStageBytes);
  __shared__ __align__(8) PipeBars bars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t0, t1;
  cta_range(p.g.num_tiles, t0, t1);          // k-stages: DW_BI x BJ blocks of pair rows

  if (threadIdx.x == 0) init_bars<kDwStages>(bars, kProducerWarps, kEpiThreads);
  if (warp == kMmaWarp) tmem_alloc(smem_u32(&bars.tmem_holder), C::kTmemCols);
  if (warp == kTmaWarp && lane == 0) tma_prefetch_desc(&map_y2);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_holder;

  if (warp < kProducerWarps) {
    // all 8 warps produce both operands of a k-stage (32 pair rows x H): thread = (float4 column c16, row residue
    // rbase), rows r = rbase + kRowStep i.  A warp covers whole (row, column half) pairs, so the raw Y2 bytes it reads
    // are exactly the ones it overwrites with the dY2 operand (in place); H1 is generated into the second region.
    const int tid = threadIdx.x;
    const int c16 = tid % C::F4, rbase = tid / C::F4;
    const int c = c16 * 4, cb = c16 >> 4, c4 = c16 & 15;
    constexpr int JJ = BJ / C::kRowStep;                  // distinct jl per thread; il = i / JJ
    const Dy2Consts k2 = load_dy2_consts(p.bn2, p.w3, p.s1, p.s2, p.inv_count, c);
    const H1Consts k1 = load_h1_consts(p.bn1, c);
    uint32_t it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
      const int s = it % kDwStages;
      const uint32_t ph = (it / kDwStages) & 1;
      uint8_t* op = smem + s * C::kStageBytes;
      // raw box x (32 columns) sits at x * 4 KB: even boxes alias the hi plane of sub-tile x / 2, odd boxes its lo plane
      const uint8_t* raw = op + (2 * cb + (c4 >> 3)) * (DW_ROWS * 128) + (c4 & 7) * 16;
      uint8_t* hi_a = op + cb * C::kSub;
      uint8_t* hi_b = hi_a + C::kOperand;
      float de[C::kIters];
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < C::kIters; ++i) {
        const int r = rbase + C::kRowStep * i;
        const int li = ib * DW_BI + (r >> 3), j = jb * BJ + (r & 7);
        const bool valid = li < p.g.n_i && j < p.g.n;
        de[i] = valid ? __ldg(p.dE + (int64_t)li * p.g.n + j) : 0.f;
        vmask |= (valid ? 1u : 0u) << i;
      }
      float4 pb[DW_BI], pa[JJ];
#pragma unroll
      for (int il = 0; il < DW_BI; ++il) {
        const int li = ib * DW_BI + il;
        pb[il] = ld4(p.Pb + (int64_t)(p.g.i_first + (li < p.g.n_i ? li : 0)) * H + c);
      }
#pragma unroll
      for (int jj = 0; jj < JJ; ++jj) {
        const int j = jb * BJ + rbase + C::kRowStep * jj;
        pa[jj] = ld4(p.Pa + (int64_t)(j < p.g.n ? j : 0) * H + c);
      }
      mbar_wait(smem_u32(&bars.raw[s]), ph);       // the TMA was issued after the stage was released: both regions are free
      {
        float4 y[C::kIters];
#pragma unroll
        for (int i = 0; i < C::kIters; ++i)
          y[i] = *reinterpret_cast<const float4*>(raw + (rbase + C::kRowStep * i) * 128);
        __syncwarp();            // reads complete before the overwrite
#pragma unroll
        for (int i = 0; i < C::kIters; ++i) {
          const int r = rbase + C::kRowStep * i;
          float4 o = dy2_value4(y[i], de[i], k2);
          if (!((vmask >> i) & 1u)) o = make_float4(0.f, 0.f, 0.f, 0.f);
          store_split4<C::kWithLo>(hi_a, hi_a + DW_ROWS * 128, sw128_offset(r, c4 * 4), o);
        }
      }
#pragma unroll
      for (int i = 0; i < C::kIters; ++i) {
        const int r = rbase + C::kRowStep * i;
        float4 o = h1_value4(pa[i % JJ], pb[i / JJ], k1);
        if (!((vmask >> i) & 1u)) o = make_float4(0.f, 0.f, 0.f, 0.f);
        store_split4<C::kWithLo>(hi_b, hi_b + DW_ROWS * 128, sw128_offset(r, c4 * 4), o);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars.full[s]));
    }
  } else if (warp == kTmaWarp) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const int ib = t / p.g.tiles_j, jb = t - ib * p.g.tiles_j;
        const int s = it % kDwStages;
        const uint32_t ph = (it / kDwStages) & 1;
        mbar_wait(smem_u32(&bars.empty[s]), ph ^ 1);
        const uint32_t dst = smem_u32(smem + s * C::kStageBytes);
        const uint32_t braw = smem_u32(&bars.raw[s]);
        mbar_arrive_expect_tx(braw, C::kOperand);
#pragma unroll
        for (int x = 0; x < H / 32; ++x)
          tma_load_3d(dst + x * (DW_ROWS * 128), &map_y2, x * 32, jb * BJ, ib * DW_BI, braw);
      }
    }
  } else if (warp == kMmaWarp) {
    uint32_t it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int s = it % kDwStages;
      const uint32_t ph = (it / kDwStages) & 1;
      mbar_wait(smem_u32(&bars.full[s]), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + s * C::kStageBytes);
        const uint32_t sb = sa + C::kOperand;
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          const uint32_t a_plane = (pass == 2) ? DW_ROWS * 128 : 0;
          const uint32_t b_plane = (pass == 1) ? DW_ROWS * 128 : 0;
#pragma unroll
          for (int ks = 0; ks < DW_ROWS / 16; ++ks) {
            const uint64_t bd = make_desc_mn_sw128(sb + b_plane + ks * 2048, C::kSub);
#pragma unroll
            for (int mh = 0; mh < C::MH; ++mh) {
              const uint64_t ad = make_desc_mn_sw128(sa + mh * 2 * C::kSub + a_plane + ks * 2048, C::kSub);
              umma_bf16(tmem_base + (uint32_t)(mh * H), ad, bd, C::kIdesc, (it | (uint32_t)pass | (uint32_t)ks) ? 1u : 0u);
            }
          }
        }
        umma_commit(smem_u32(&bars.empty[s]));
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(smem_u32(&bars.tfull[0]));
    __syncwarp();
  } else {
    // epilogue: the CTA's h x h partial sum -> global atomics (coalesced float4 rows through the staging tile)
    const int q = warp & 3;
    float* st = staging + q * (32 * 36);
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    mbar_wait(smem_u32(&bars.tfull[0]), 0);
    tc_fence_after();
#pragma unroll 1
    for (int mh = 0; mh < C::MH; ++mh) {
#pragma unroll 1
      for (int b = 0; b < H / 32; ++b) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * H + b * 32), r);
        tmem_ld_wait();
        stage_block(st, lane, r);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int row = mh * 128 + q * 32 + rsub + 4 * k;
          atomicAdd(reinterpret_cast<float4*>(p.dW2 + (int64_t)row * H + b * 32 + c4),
                    *reinterpret_cast<const float4*>(st + (rsub + 4 * k) * 36 + c4));
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ small kernels
__global__ void stats_finalize_kernel(int h, const double* __restrict__ stats, double count, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= h) return;
  const double m = stats[k] / count;
  double var = stats[h + k] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[k] = (float)m;
  rstd[k] = (float)(1.0 / sqrt(var + (double)eps));
}

// BN1 backward sums from the reduced tile gradients (work = [t1 | t2 (2h doubles) | Ga (n x h) | Gb (n x h)]):
//   t1[c] = sum_j Ga[j,c],   t2[c] = rstd[c] * (sum_j Ga[j,c] (Pa[j,c] - mean_a[c]) + sum_i Gb[i,c] (Pb[i,c] - mean_b[c]))
// because xhat[(i,j),c] = rstd[c] ((Pa[j,c] - mean_a[c]) + (Pb[i,c] - mean_b[c])).  Adds into t1/t2 (partial Gb of a
// rank's slice gives that rank's partial sums).
__global__ void __launch_bounds__(256)
bn1_tsum_kernel(int n, int h, const float* __restrict__ Pa, const float* __restrict__ Pb,
                const float* __restrict__ col_mean, const float* __restrict__ rstd, const float* __restrict__ Ga,
                const float* __restrict__ Gb, double* __restrict__ tsum) {
  // 32 columns x 8 row groups per block; blockIdx.y splits the rows (one double atomic per column, statistic and block)
  __shared__ double sm[2][8][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int per = (n + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(n, r0 + per);
  double a1 = 0.0, a2 = 0.0;
  if (c < h) {
    const double ma = (double)col_mean[c], mb = (double)col_mean[h + c];
    for (int r = r0 + rg; r < r1; r += 8) {
      const double ga = (double)Ga[(int64_t)r * h + c], gb = (double)Gb[(int64_t)r * h + c];
      a1 += ga;
      a2 += ga * ((double)Pa[(int64_t)r * h + c] - ma) + gb * ((double)Pb[(int64_t)r * h + c] - mb);
    }
  }
  sm[0][rg][cl] = a1;
  sm[1][rg][cl] = a2;
  __syncthreads();
  if (rg == 0 && c < h) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      t1 += sm[0][g][cl];
      t2 += sm[1][g][cl];
    }
    atomicAdd(tsum + c, t1);
    atomicAdd(tsum + h + c, t2 * (double)rstd[c]);
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// Y2 slice (n_i * n rows x h fp32) as a 3-D tensor (column, j, i); a box is 32 columns x BJ j x box_i i
static int encode_y2_map(CUtensorMap* map, const float* Y2, int h, int n, int n_i, int box_i, bool swizzle128 = false) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error_msg("cuTensorMapEncodeTiled is not available from this driver");
    return GS_ENOSYS;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)h, (cuuint64_t)n, (cuuint64_t)n_i};
  const cuuint64_t strides[2] = {(cuuint64_t)h * 4, (cuuint64_t)n * h * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)BJ, (cuuint32_t)box_i};
  const cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(Y2), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    std::snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    set_error_msg(msg);
    return GS_EINVAL;
  }
  return GS_OK;
}

static Geom make_geom(int n, int n_i, int i_first, int bi) {
  Geom g{n, n_i, i_first, (n + BJ - 1) / BJ, 0};
  g.num_tiles = g.tiles_j * ((n_i + bi - 1) / bi);
  return g;
}

template <typename KernelT>
static int set_smem(KernelT kernel, size_t bytes, const char* what) {
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error(what, e);
    return (int)e;
  }
  return GS_OK;
}

static int pack_image(const float* W, int64_t ldw, int h, int trans, int planes, void* ws, cudaStream_t st) {
  pack_w_image_kernel<<<(h * h + 255) / 256, 256, 0, st>>>(W, ldw, h, trans, planes, reinterpret_cast<uint8_t*>(ws));
  return finish_launch("pack_w_image");
}

template <int H, int NPASS>
static int launch_fwd(FwdParams& p, cudaStream_t st) {
  using C = Cfg<H, NPASS>;
  constexpr size_t smem = (size_t)C::kStages * C::kStageBytes + kFwdStagingBytes + 1024;
  static bool configured = false;
  if (!configured) {
    int rc = set_smem(pge_l2_fwd_kernel<H, NPASS, true>, smem, "cudaFuncSetAttribute(pge_l2_fwd)");
    if (rc) return rc;
    rc = set_smem(pge_l2_fwd_kernel<H, NPASS, false>, smem, "cudaFuncSetAttribute(pge_l2_fwd)");
    if (rc) return rc;
    configured = true;
  }
  // GS_PGE_TMA_STORE=0: the epilogue stores Y2 with per-lane float4 stores instead of one TMA store per block
  static const int tma_store = [] {
    const char* e = getenv("GS_PGE_TMA_STORE");
    return e ? atoi(e) : 1;
  }();
  static const int two_cta = [] {
    const char* e = getenv("GS_PGE_2CTA");
    return e ? atoi(e) : 0;
  }();
  if (two_cta && p.g.num_tiles >= 2 * kNumSMs) {
    constexpr size_t smem2 = (size_t)C::KB * C::kPlanes * (H / 2) * BK * 2 + 2 * kStageA + kFwdStagingBytes + 1024;
    static bool configured2 = false;
    if (!configured2) {
      const int rc = set_smem(pge_l2_fwd2_kernel<H, NPASS>, smem2, "cudaFuncSetAttribute(pge_l2_fwd2)");
      if (rc) return rc;
      configured2 = true;
    }
    pge_l2_fwd2_kernel<H, NPASS><<<kNumSMs, kThreads, smem2, st>>>(p);
    return finish_launch("pge_l2_fwd2");
  }
  const int grid = p.g.num_tiles < kNumSMs ? p.g.num_tiles : kNumSMs;
  CUtensorMap map;
  const int rc = encode_y2_map(&map, p.Y2, H, p.g.n, p.g.n_i, 4, true);   // box: 32 columns x 8 j x 4 i, SWIZZLE_128B
  if (rc) return rc;
  if (tma_store) pge_l2_fwd_kernel<H, NPASS, true><<<grid, kFwdThreads, smem, st>>>(map, p);
  else pge_l2_fwd_kernel<H, NPASS, false><<<grid, kFwdThreads, smem, st>>>(map, p);
  return finish_launch("pge_l2_fwd");
}

template <int H, int NPASS, bool kFused>
static int launch_dx(const CUtensorMap& map, BwdParams& p, cudaStream_t st) {
  using C = Cfg<H, NPASS>;
  constexpr size_t smem = (size_t)C::kStages * C::kStageBytes + kStagingBytes + kGaBytes + 1024;
  static bool configured = false;
  if (!configured) {
    const int rc = set_smem(pge_l2_bwd_dx_kernel<H, NPASS, kFused>, smem, "cudaFuncSetAttribute(pge_l2_bwd_dx)");
    if (rc) return rc;
    configured = true;
  }
  const int grid = p.g.num_tiles < kNumSMs ? p.g.num_tiles : kNumSMs;
  pge_l2_bwd_dx_kernel<H, NPASS, kFused><<<grid, kThreads, smem, st>>>(map, p);
  return finish_launch("pge_l2_bwd_dx");
}

template <int H, int NPASS>
static int launch_dw(const CUtensorMap& map, BwdParams& p, cudaStream_t st) {
  using C = DwCfg<H, NPASS>;
  constexpr size_t smem = (size_t)kDwStages * C::kStageBytes + kStagingBytes + 1024;
  static bool configured = false;
  if (!configured) {
    const int rc = set_smem(pge_l2_bwd_dw_kernel<H, NPASS>, smem, "cudaFuncSetAttribute(pge_l2_bwd_dw)");
    if (rc) return rc;
    configured = true;
  }
  const int grid = p.g.num_tiles < kNumSMs ? p.g.num_tiles : kNumSMs;
  pge_l2_bwd_dw_kernel<H, NPASS><<<grid, kThreads, smem, st>>>(map, p);
  return finish_launch("pge_l2_bwd_dw");
}

}  // namespace pf
}  // namespace gs

extern "C" {
using namespace gs;

#define GS_PF_REQ_SHAPE                                                                                      \
  GS_REQUIRE(n > 0 && n_i > 0 && i_first >= 0 && i_first + n_i <= n && (h == 128 || h == 256) &&            \
             (precision == 1 || precision == 2))

int64_t gs_pge_fused_workspace_bytes(int32_t h, int precision) {
  return (int64_t)h * h * 2 * (precision == 1 ? 2 : 1);
}

int gs_pge_fused_l2_fwd_f32(int32_t n, int32_t n_i, int32_t i_first, int32_t h, const float* Pa, const float* Pb,
                            const float* mean1, const float* rstd1, const float* gamma1, const float* beta1,
                            const float* W2, int64_t ldw, float* Y2, double* stats, int precision, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  GS_PF_REQ_SHAPE;
  GS_REQUIRE(Pa && Pb && mean1 && rstd1 && gamma1 && beta1 && W2 && Y2 && stats && workspace && ldw >= h);
  GS_REQUIRE(workspace_bytes >= gs_pge_fused_workspace_bytes(h, precision));
  GS_REQUIRE(((reinterpret_cast<uintptr_t>(Pa) | reinterpret_cast<uintptr_t>(Pb) | reinterpret_cast<uintptr_t>(Y2) |
               reinterpret_cast<uintptr_t>(workspace)) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  const int planes = precision == 1 ? 2 : 1;
  int rc = pf::pack_image(W2, ldw, h, 0, planes, workspace, st);
  if (rc) return rc;
  cudaMemsetAsync(stats, 0, sizeof(double) * 2 * h, st);
  pf::FwdParams p{pf::make_geom(n, n_i, i_first, pf::BI), Pa, Pb, {mean1, rstd1, gamma1, beta1},
                  reinterpret_cast<const uint8_t*>(workspace), Y2, stats};
  if (h == 256) return precision == 1 ? pf::launch_fwd<256, 3>(p, st) : pf::launch_fwd<256, 1>(p, st);
  return precision == 1 ? pf::launch_fwd<128, 3>(p, st) : pf::launch_fwd<128, 1>(p, st);
}

int gs_pge_stats_finalize_f32(int32_t h, const double* stats, double count, float eps, float* mean, float* rstd,
                              void* stream) {
  GS_REQUIRE(h > 0 && stats && count > 0 && mean && rstd);
  pf::stats_finalize_kernel<<<(h + 255) / 256, 256, 0, as_stream(stream)>>>(h, stats, count, eps, mean, rstd);
  return finish_launch("pge_stats_finalize");
}

// dH1 = dY2 W2 with dY2 computed on the fly.  dH1 != NULL: the product is stored (unmasked); otherwise it is masked
// with H1 > 0 and reduced into Ga / Gb (n x h each, accumulated: the caller zeroes them).
int gs_pge_fused_l2_bwd_dx_f32(int32_t n, int32_t n_i, int32_t i_first, int32_t h, const float* Pa, const float* Pb,
                               const float* mean1, const float* rstd1, const float* gamma1, const float* beta1,
                               const float* W2, int64_t ldw, const float* Y2, const float* dE, const float* mean2,
                               const float* rstd2, const float* gamma2, const float* beta2, const float* w3,
                               const float* s1, const float* s2, double count, float* Ga, float* Gb, float* dH1,
                               int precision, void* workspace, int64_t workspace_bytes, void* stream) {
  GS_PF_REQ_SHAPE;
  GS_REQUIRE(Pa && Pb && mean1 && rstd1 && gamma1 && beta1 && W2 && Y2 && dE && mean2 && rstd2 && gamma2 && beta2 &&
             w3 && s1 && s2 && count > 0 && workspace && ldw >= h && (dH1 || (Ga && Gb)));
  GS_REQUIRE(workspace_bytes >= gs_pge_fused_workspace_bytes(h, precision));
  GS_REQUIRE(((reinterpret_cast<uintptr_t>(Pa) | reinterpret_cast<uintptr_t>(Pb) | reinterpret_cast<uintptr_t>(Y2) |
               reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(dH1)) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  const int planes = precision == 1 ? 2 : 1;
  int rc = pf::pack_image(W2, ldw, h, 1, planes, workspace, st);
  if (rc) return rc;
  CUtensorMap map;
  rc = pf::encode_y2_map(&map, Y2, h, n, n_i, pf::BI);
  if (rc) return rc;
  pf::BwdParams p{pf::make_geom(n, n_i, i_first, pf::BI), Pa, Pb, {mean1, rstd1, gamma1, beta1},
                  {mean2, rstd2, gamma2, beta2}, dE, w3, s1, s2, (float)(1.0 / count),
                  reinterpret_cast<const uint8_t*>(workspace), Ga, Gb, dH1, nullptr};
  const bool three = precision == 1;
  if (dH1) {
    if (h == 256) return three ? pf::launch_dx<256, 3, false>(map, p, st) : pf::launch_dx<256, 1, false>(map, p, st);
    return three ? pf::launch_dx<128, 3, false>(map, p, st) : pf::launch_dx<128, 1, false>(map, p, st);
  }
  if (h == 256) return three ? pf::launch_dx<256, 3, true>(map, p, st) : pf::launch_dx<256, 1, true>(map, p, st);
  return three ? pf::launch_dx<128, 3, true>(map, p, st) : pf::launch_dx<128, 1, true>(map, p, st);
}

// dW2 (h x h, row = output unit) = dY2^T H1 over the slice's pair rows; overwritten.
int gs_pge_fused_l2_bwd_dw_f32(int32_t n, int32_t n_i, int32_t i_first, int32_t h, const float* Pa, const float* Pb,
                               const float* mean1, const float* rstd1, const float* gamma1, const float* beta1,
                               const float* Y2, const float* dE, const float* mean2, const float* rstd2,
                               const float* gamma2, const float* beta2, const float* w3, const float* s1,
                               const float* s2, double count, float* dW2, int precision, void* stream) {
  GS_PF_REQ_SHAPE;
  GS_REQUIRE(Pa && Pb && mean1 && rstd1 && gamma1 && beta1 && Y2 && dE && mean2 && rstd2 && gamma2 && beta2 && w3 &&
             s1 && s2 && count > 0 && dW2);
  GS_REQUIRE(((reinterpret_cast<uintptr_t>(Pa) | reinterpret_cast<uintptr_t>(Pb) | reinterpret_cast<uintptr_t>(Y2) |
               reinterpret_cast<uintptr_t>(dW2)) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  CUtensorMap map;
  int rc = pf::encode_y2_map(&map, Y2, h, n, n_i, pf::DW_BI);
  if (rc) return rc;
  cudaMemsetAsync(dW2, 0, sizeof(float) * h * h, st);
  pf::BwdParams p{pf::make_geom(n, n_i, i_first, pf::DW_BI), Pa, Pb, {mean1, rstd1, gamma1, beta1},
                  {mean2, rstd2, gamma2, beta2}, dE, w3, s1, s2, (float)(1.0 / count), nullptr, nullptr, nullptr,
                  nullptr, dW2};
  const bool three = precision == 1;
  if (h == 256) return three ? pf::launch_dw<256, 3>(map, p, st) : pf::launch_dw<256, 1>(map, p, st);
  return three ? pf::launch_dw<128, 3>(map, p, st) : pf::launch_dw<128, 1>(map, p, st);
}

// work = [t1 | t2 (2h doubles) | Ga (n x h floats) | Gb (n x h floats)] (the layout of gs_pge_bn1_bwd_pass_rows_f32):
// adds the BN1 backward sums implied by Ga / Gb into t1 / t2.
int gs_pge_bn1_tsum_f64(int32_t n, int32_t h, const float* Pa, const float* Pb, const float* col_mean,
                        const float* rstd1, void* work, void* stream) {
  GS_REQUIRE(n > 0 && h > 0 && Pa && Pb && col_mean && rstd1 && work);
  double* tsum = reinterpret_cast<double*>(work);
  const float* Ga = reinterpret_cast<const float*>(tsum + 2 * h);
  const float* Gb = Ga + (int64_t)n * h;
  const int splits = n >= 512 ? 16 : (n >= 64 ? 4 : 1);
  pf::bn1_tsum_kernel<<<dim3((h + 31) / 32, splits), 256, 0, as_stream(stream)>>>(n, h, Pa, Pb, col_mean, rstd1, Ga, Gb, tsum);
  return finish_launch("pge_bn1_tsum");
}

}  // extern "C"
