// Grouped K-segmented products C_g = A[rows of g]^T B[rows of g] on tcgen05 with BOTH operands MN-major and TMA-fed
// (sm_100a) -- the per-class weight gradients of the real side (autograd.grad(loss_real, params) of
// condensation/gcond_base.py:221-224 for all classes at once).
//
// A (rows x M) and B (rows x N) are row-major fp32 matrices whose rows are grouped by class (64-aligned segments, the
// sampler's padding).  A row-major [k row][column] block IS the MN-major UMMA operand layout (tc_ptx.cuh), so nothing is
// transposed or packed: TMA (cp.async.bulk.tensor.2d) drops 32-row x 32-column fp32 boxes into the operand stage where
// the hi plane of a 64-column sub-tile aliases the first box and the lo plane the second, and the producer warps split
// them to BF16 hi/lo IN PLACE.  One group's M x N result accumulates in TMEM over the group's rows and is flushed with
// float4 atomics when the CTA's (contiguous) share of k-stages moves to the next group.
// This replaces gs_gemm_grouped_tn_f32's path of pack_b (B re-written as a BF16 image in HBM) + transposing producers.
//
// B modes:
//   0  B is a wide matrix (N = 64 * BSUB columns), TMA boxes converted in place like A;
//   1  B is narrow (N <= 64, N % 4 == 0: class-width gradients): one TMA box per stage into a side buffer, converted into
//      a single 64-column sub-tile whose unused columns stay zero;
//   2  B is GENERATED: dA1 = (dU W2^T) . [H1 > 0] -- the backward of the hidden ReLU layer of the condense model
//      (models/sgc.py:37-57, models/layers.py:36-51 under autograd) computed per stage on the CUDA cores from the narrow
//      dU box (side buffer), W2^T kept in shared memory and the mask rows of H1; its column sums per group (the bias
//      gradient) come out on the way.  dA1 never exists in HBM.
#include <cstdio>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gs {
namespace gt {

using namespace gs::ptx;

constexpr int KR = 32;                 // rows per k-stage
constexpr int kMaxStages = 3;
constexpr int kProducerWarps = 8, kProducerThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kMmaWarp = 12, kTmaWarp = 13;
constexpr int kThreads = 14 * 32;
constexpr uint32_t kSub = 2 * KR * 128;             // one 64-column sub-tile: hi plane (32 rows x 128 B) | lo plane
constexpr uint32_t kStagingBytes = 4 * 32 * 36 * 4;
constexpr int kMaxCw = 64;                          // widest narrow matrix (side buffer)
constexpr uint32_t kSideBytes = KR * kMaxCw * 4;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct Params {
  int total_rows, stages;
  int G;
  const int32_t* seg;          // G + 1 row offsets (multiples of 64)
  const int32_t* out_block;    // column block of C per group
  int M, N;                    // logical output size per group
  float* C;                    // M x (nblk * N), zeroed by the caller
  int64_t ldc;
  // mode 2
  const float* H1;             // rows x N mask source
  int64_t ldh;
  const float* W2;             // N x Cw:  dA1[r, c] = sum_k dU[r, k] W2[c, k]
  int64_t ldw2;
  int Cw;                      // columns of the narrow matrix (modes 1, 2)
  float* gb;                   // 1 x (nblk * N): per-group column sums of the generated B (mode 2)
};

struct Bars {
  uint64_t raw[kMaxStages], full[kMaxStages], empty[kMaxStages], tfull, tempty;
  uint32_t tmem_holder;
};

template <bool kWithLo>
__device__ __forceinline__ void store_split4(uint8_t* hi_plane, uint32_t off, const float4& v) {
  uint2 ph, pl;
  split_bf16x2(v.x, v.y, ph.x, pl.x);
  split_bf16x2(v.z, v.w, ph.y, pl.y);
  *reinterpret_cast<uint2*>(hi_plane + off) = ph;
  if (kWithLo) *reinterpret_cast<uint2*>(hi_plane + KR * 128 + off) = pl;
}

// raw fp32 boxes of a (KR x 64*SUBS) operand region -> BF16 hi/lo MN-major sub-tiles, in place.  Thread = (float4
// column c16, row residue); a warp covers whole (row, sub-tile) pairs, so what it reads is what it overwrites.
template <int SUBS, bool kWithLo>
__device__ __forceinline__ void convert_in_place(uint8_t* region, int tid) {
  constexpr int F4 = 16 * SUBS;
  constexpr int kRowStep = kProducerThreads / F4;
  constexpr int kIters = KR / kRowStep;
  const int c16 = tid % F4, rbase = tid / F4;
  const int cb = c16 >> 4, c4 = c16 & 15;
  const uint8_t* raw = region + (2 * cb + (c4 >> 3)) * (KR * 128) + (c4 & 7) * 16;
  uint8_t* hi = region + cb * kSub;
  float4 y[kIters];
#pragma unroll
  for (int i = 0; i < kIters; ++i) y[i] = *reinterpret_cast<const float4*>(raw + (rbase + kRowStep * i) * 128);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < kIters; ++i) store_split4<kWithLo>(hi, sw128_offset(rbase + kRowStep * i, c4 * 4), y[i]);
}

template <int ASUB, int BSUB, int NPASS, int BMODE>
struct Cfg {
  static constexpr bool kWithLo = NPASS == 3;
  static constexpr int kStages = BMODE == 2 ? 2 : 3;                // mode 2 keeps W2^T (up to 64 KB) in shared memory
  static constexpr uint32_t kA = ASUB * kSub, kB = BSUB * kSub;
  static constexpr uint32_t kSide = BMODE == 0 ? 0 : kSideBytes;
  static constexpr uint32_t kStageBytes = kA + kB + kSide;
  static constexpr int MH = ASUB / 2;                               // 128-row halves of the result
  static constexpr int kNmmaWide = 64 * BSUB;
  static constexpr uint32_t kAccStride = 64 * BSUB;                 // TMEM columns reserved per half
  static constexpr uint32_t kTmemNeed = MH * kAccStride;
  static constexpr uint32_t kTmemCols = kTmemNeed <= 32 ? 32 : (kTmemNeed <= 64 ? 64 : (kTmemNeed <= 128 ? 128 : (kTmemNeed <= 256 ? 256 : 512)));
  static_assert(ASUB == 2 || ASUB == 4, "A must be 128 or 256 columns wide");
  static_assert(kTmemNeed <= 512, "result does not fit TMEM");
};

// group of the k-stage starting at `row0`, advancing monotonically from g (empty groups are skipped)
__device__ __forceinline__ int group_of(const Params& p, int row0, int g) {
  while (g < p.G && row0 >= __ldg(p.seg + g + 1)) ++g;
  return g;
}

template <int ASUB, int BSUB, int NPASS, int BMODE>
__global__ void __launch_bounds__(kThreads, 1)
grouped_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Params p) {
  using C = Cfg<ASUB, BSUB, NPASS, BMODE>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(smem + kStages * C::kStageBytes);
  float* w2t = reinterpret_cast<float*>(smem + kStages * C::kStageBytes + kStagingBytes);     // mode 2: [Cw][64 * BSUB]
  __shared__ __align__(8) Bars bars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = (int)(((int64_t)p.stages * blockIdx.x) / gridDim.x);
  const int s1 = (int)(((int64_t)p.stages * (blockIdx.x + 1)) / gridDim.x);
  const int n_mma = BMODE == 1 ? ((p.N + 15) & ~15) : C::kNmmaWide;
  const uint32_t idesc = make_idesc_bf16(128, n_mma, 1, 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bars.raw[s]), 1);
      mbar_init(smem_u32(&bars.full[s]), kProducerWarps);
      mbar_init(smem_u32(&bars.empty[s]), 1);
    }
    mbar_init(smem_u32(&bars.tfull), 1);
    mbar_init(smem_u32(&bars.tempty), kEpiThreads);
    fence_barrier_init();
  }
  if (BMODE == 1) {
    // the narrow operand fills only N of its sub-tile's 64 columns: the rest must read as zero
    for (int s = 0; s < kStages; ++s) {
      uint4* bz = reinterpret_cast<uint4*>(smem + s * C::kStageBytes + C::kA);
      for (int i = threadIdx.x; i < (int)(C::kB / 16); i += kThreads) bz[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  if (BMODE == 2) {
    for (int i = threadIdx.x; i < p.Cw * p.N; i += kThreads) {
      const int c = i / p.Cw, k = i % p.Cw;                       // W2 is N x Cw row-major
      w2t[k * (64 * BSUB) + c] = __ldg(p.W2 + (int64_t)c * p.ldw2 + k);
    }
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(&bars.tmem_holder), C::kTmemCols);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_holder;

  if (warp < kProducerWarps) {
    // ============================== producers ==============================
    const int tid = threadIdx.x;
    int g = 0;
    int cur_g = -1;
    float4 gbsum = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int F4B = 16 * BSUB;
    const int bc16 = tid % F4B, brbase = tid / F4B;               // mode 2 mapping over the generated columns
    auto flush_gb = [&]() {
      if (BMODE == 2 && cur_g >= 0 && bc16 * 4 < p.N)
        atomicAdd(reinterpret_cast<float4*>(p.gb + (int64_t)__ldg(p.out_block + cur_g) * p.N + bc16 * 4), gbsum);
      gbsum = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    uint32_t it = 0;
    for (int st = s0; st < s1; ++st, ++it) {
      const int row0 = st * KR;
      g = group_of(p, row0, g);
      if (g >= p.G) break;                                        // rows past the last segment (none by contract)
      if (g != cur_g) {
        flush_gb();
        cur_g = g;
      }
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      uint8_t* region_a = smem + s * C::kStageBytes;
      uint8_t* region_b = region_a + C::kA;
      mbar_wait(smem_u32(&bars.raw[s]), ph);
      convert_in_place<ASUB, C::kWithLo>(region_a, tid);
      if (BMODE == 0) {
        convert_in_place<BSUB, C::kWithLo>(region_b, tid);
      } else if (BMODE == 1) {
        const float* side = reinterpret_cast<const float*>(region_b + C::kB);
        const int f4 = p.Cw >> 2;                                 // float4 per row of the narrow matrix
        for (int i = tid; i < KR * f4; i += kProducerThreads) {
          const int r = i / f4, q = i % f4;
          store_split4<C::kWithLo>(region_b, sw128_offset(r, q * 4), *reinterpret_cast<const float4*>(side + r * p.Cw + q * 4));
        }
      } else {
        // dA1[r, c] = (sum_k dU[r, k] W2[c, k]) * [H1[r, c] > 0] for this thread's 4 columns and KR / rowstep rows
        constexpr int kRowStep = kProducerThreads / F4B;
        constexpr int kIters = KR / kRowStep;
        const float* side = reinterpret_cast<const float*>(region_b + C::kB);
        const int c = bc16 * 4, cb = bc16 >> 4, c4 = bc16 & 15;
        float4 h1[kIters];
        float2 acc_lo[kIters], acc_hi[kIters];          // packed pairs: one FFMA2 (fma.rn.f32x2) per two columns
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const int row = row0 + brbase + kRowStep * i;
          h1[i] = (row < p.total_rows && c < p.N) ? ld4(p.H1 + (int64_t)row * p.ldh + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          acc_lo[i] = acc_hi[i] = make_float2(0.f, 0.f);
        }
        for (int k = 0; k < p.Cw; ++k) {
          const float4 w = *reinterpret_cast<const float4*>(w2t + k * (64 * BSUB) + c);
          const float2 w_lo = make_float2(w.x, w.y), w_hi = make_float2(w.z, w.w);
#pragma unroll
          for (int i = 0; i < kIters; ++i) {
            const float du = side[(brbase + kRowStep * i) * p.Cw + k];
            const float2 du2 = make_float2(du, du);
            acc_lo[i] = __ffma2_rn(du2, w_lo, acc_lo[i]);
            acc_hi[i] = __ffma2_rn(du2, w_hi, acc_hi[i]);
          }
        }
        uint8_t* hi = region_b + cb * kSub;
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          float4 o;
          o.x = h1[i].x > 0.f ? acc_lo[i].x : 0.f;
          o.y = h1[i].y > 0.f ? acc_lo[i].y : 0.f;
          o.z = h1[i].z > 0.f ? acc_hi[i].x : 0.f;
          o.w = h1[i].w > 0.f ? acc_hi[i].y : 0.f;
          gbsum.x += o.x; gbsum.y += o.y; gbsum.z += o.z; gbsum.w += o.w;
          store_split4<C::kWithLo>(hi, sw128_offset(brbase + kRowStep * i, c4 * 4), o);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars.full[s]));
    }
    flush_gb();
  } else if (warp == kTmaWarp) {
    if (lane == 0) {
      int g = 0;
      uint32_t it = 0;
      for (int st = s0; st < s1; ++st, ++it) {
        const int row0 = st * KR;
        g = group_of(p, row0, g);
        if (g >= p.G) break;
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(smem_u32(&bars.empty[s]), ph ^ 1);
        const uint32_t dst = smem_u32(smem + s * C::kStageBytes);
        const uint32_t braw = smem_u32(&bars.raw[s]);
        const uint32_t side_bytes = BMODE == 0 ? 0u : (uint32_t)(KR * p.Cw * 4);
        mbar_arrive_expect_tx(braw, C::kA + (BMODE == 0 ? C::kB : 0u) + side_bytes);
#pragma unroll
        for (int x = 0; x < 2 * ASUB; ++x) tma_load_2d(dst + x * (KR * 128), &map_a, x * 32, row0, braw);
        if (BMODE == 0) {
#pragma unroll
          for (int x = 0; x < 2 * BSUB; ++x) tma_load_2d(dst + C::kA + x * (KR * 128), &map_b, x * 32, row0, braw);
        } else {
          tma_load_2d(dst + C::kA + C::kB, &map_b, 0, row0, braw);          // narrow box: Cw columns x KR rows
        }
      }
    }
  } else if (warp == kMmaWarp) {
    int g = 0, cur_g = -1;
    uint32_t it = 0, flushes = 0;
    bool first = true;
    for (int st = s0; st < s1; ++st, ++it) {
      const int row0 = st * KR;
      g = group_of(p, row0, g);
      if (g >= p.G) break;
      if (g != cur_g) {
        if (cur_g >= 0) {
          if (lane == 0) umma_commit(smem_u32(&bars.tfull));         // the finished group's accumulator is complete
          __syncwarp();
          mbar_wait(smem_u32(&bars.tempty), flushes & 1);            // ... and drained before it is overwritten
          tc_fence_after();
          ++flushes;
        }
        cur_g = g;
        first = true;
      }
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(smem_u32(&bars.full[s]), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + s * C::kStageBytes);
        const uint32_t sb = sa + C::kA;
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          const uint32_t a_plane = (pass == 2) ? KR * 128 : 0;
          const uint32_t b_plane = (pass == 1) ? KR * 128 : 0;
#pragma unroll
          for (int ks = 0; ks < KR / 16; ++ks) {
            const uint64_t bd = make_desc_mn_sw128(sb + b_plane + ks * 2048, kSub);
#pragma unroll
            for (int mh = 0; mh < C::MH; ++mh) {
              const uint64_t ad = make_desc_mn_sw128(sa + mh * 2 * kSub + a_plane + ks * 2048, kSub);
              umma_bf16(tmem_base + (uint32_t)(mh * C::kAccStride), ad, bd, idesc,
                        (first && pass == 0 && ks == 0) ? 0u : 1u);
            }
          }
        }
        umma_commit(smem_u32(&bars.empty[s]));
      }
      __syncwarp();
      first = false;
    }
    if (cur_g >= 0) {
      if (lane == 0) umma_commit(smem_u32(&bars.tfull));
      __syncwarp();
    }
  } else {
    // ============================== epilogue: one flush per group this CTA touched ==============================
    const int q = warp & 3;
    float* stg = staging + q * (32 * 36);
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    int g = 0, cur_g = -1;
    uint32_t flushes = 0;
    auto flush = [&](int gf) {
      mbar_wait(smem_u32(&bars.tfull), flushes & 1);
      tc_fence_after();
      float* cblk = p.C + (int64_t)__ldg(p.out_block + gf) * p.N;
      const int nblk32 = (p.N + 31) / 32;
#pragma unroll 1
      for (int mh = 0; mh < C::MH; ++mh) {
#pragma unroll 1
        for (int b = 0; b < nblk32; ++b) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * C::kAccStride + b * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stg + lane * 36 + j) = make_float4(
                __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = mh * 128 + q * 32 + rsub + 4 * k;
            const int colc = b * 32 + c4;
            if (row < p.M && colc < p.N)
              atomicAdd(reinterpret_cast<float4*>(cblk + (int64_t)row * p.ldc + colc),
                        *reinterpret_cast<const float4*>(stg + (rsub + 4 * k) * 36 + c4));
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars.tempty));
      ++flushes;
    };
    for (int st = s0; st < s1; ++st) {
      g = group_of(p, st * KR, g);
      if (g >= p.G) break;
      if (g != cur_g) {
        if (cur_g >= 0) flush(cur_g);
        cur_g = g;
      }
    }
    if (cur_g >= 0) flush(cur_g);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
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
// rows x cols fp32 matrix with leading dimension ld (elements); a box is box_cols columns x KR rows
static int encode_2d(CUtensorMap* map, const float* base, int64_t rows, int cols, int64_t ld, int box_cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error_msg("cuTensorMapEncodeTiled is not available from this driver");
    return GS_ENOSYS;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)KR};
  const cuuint32_t es[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    std::snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    set_error_msg(msg);
    return GS_EINVAL;
  }
  return GS_OK;
}

template <int ASUB, int BSUB, int NPASS, int BMODE>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, Params& p, cudaStream_t st) {
  using C = Cfg<ASUB, BSUB, NPASS, BMODE>;
  const size_t smem = (size_t)C::kStages * C::kStageBytes + kStagingBytes +
                      (BMODE == 2 ? (size_t)p.Cw * 64 * BSUB * 4 : 0) + 1024;
  constexpr size_t kSmemCap = 227 * 1024 - 512;              // opt-in limit minus the static barriers
  if (smem > kSmemCap) return GS_ENOSYS;
  static bool configured = false;
  if (!configured) {
    const cudaError_t e = cudaFuncSetAttribute(grouped_tn_kernel<ASUB, BSUB, NPASS, BMODE>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(grouped_tn)", e);
      return (int)e;
    }
    configured = true;
  }
  const int grid = p.stages < kNumSMs ? p.stages : kNumSMs;
  grouped_tn_kernel<ASUB, BSUB, NPASS, BMODE><<<grid, kThreads, smem, st>>>(ma, mb, p);
  return finish_launch("grouped_tn");
}

template <int ASUB, int NPASS>
static int dispatch_b(int bmode, int bsub, const CUtensorMap& ma, const CUtensorMap& mb, Params& p, cudaStream_t st) {
  if (bmode == 1) return launch<ASUB, 1, NPASS, 1>(ma, mb, p, st);
  if (bmode == 2) return bsub == 4 ? launch<ASUB, 4, NPASS, 2>(ma, mb, p, st) : GS_ENOSYS;
  if (bsub == 4) return launch<ASUB, 4, NPASS, 0>(ma, mb, p, st);
  if (bsub == 2) return launch<ASUB, 2, NPASS, 0>(ma, mb, p, st);
  return GS_ENOSYS;
}

static int dispatch(int asub, int bmode, int bsub, int precision, const CUtensorMap& ma, const CUtensorMap& mb, Params& p,
                    cudaStream_t st) {
  if (asub == 2) return precision == 1 ? dispatch_b<2, 3>(bmode, bsub, ma, mb, p, st) : dispatch_b<2, 1>(bmode, bsub, ma, mb, p, st);
  if (asub == 4) return precision == 1 ? dispatch_b<4, 3>(bmode, bsub, ma, mb, p, st) : dispatch_b<4, 1>(bmode, bsub, ma, mb, p, st);
  return GS_ENOSYS;
}

}  // namespace gt
}  // namespace gs

extern "C" {
using namespace gs;

static inline bool gt_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 1 when gs_gemm_grouped_mn_f32 / gs_mlp_bwd_grouped_f32 cover the shape (else the caller uses gs_gemm_grouped_tn_f32)
int gs_gemm_grouped_mn_supported(int32_t M, int32_t N, int precision) {
  const bool a_ok = (M == 128 || M == 256);
  const bool b_ok = (N == 128 || N == 256) || (N >= 4 && N <= gt::kMaxCw && N % 4 == 0);
  return (a_ok && b_ok && (precision == 1 || precision == 2)) ? 1 : 0;
}

// C[:, out_block[g]*N : +N] += A[seg[g]:seg[g+1]]^T B[seg[g]:seg[g+1]]  (C zeroed by the caller; segments 64-aligned)
int gs_gemm_grouped_mn_f32(int32_t G, const int32_t* seg, const int32_t* out_block, int32_t M, int32_t N,
                           int32_t total_rows, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                           int64_t ldc, int precision, void* stream) {
  GS_REQUIRE(G > 0 && seg && out_block && A && B && C && total_rows > 0 && lda >= M && ldb >= N && ldc >= N);
  GS_REQUIRE(gs_gemm_grouped_mn_supported(M, N, precision));
  GS_REQUIRE(gt_aligned16(A) && gt_aligned16(B) && gt_aligned16(C) && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0);
  const bool narrow = N <= gt::kMaxCw;
  CUtensorMap ma, mb;
  int rc = gt::encode_2d(&ma, A, total_rows, M, lda, 32);
  if (rc) return rc;
  rc = gt::encode_2d(&mb, B, total_rows, N, ldb, narrow ? N : 32);
  if (rc) return rc;
  gt::Params p{total_rows, (total_rows + gt::KR - 1) / gt::KR, G, seg, out_block, M, N, C, ldc, nullptr, 0, nullptr, 0,
               narrow ? N : 0, nullptr};
  return gt::dispatch(M / 64, narrow ? 1 : 0, narrow ? 1 : N / 64, precision, ma, mb, p, gs::as_stream(stream));
}

// Backward of the hidden layer of the 2-layer condense models on the real side, all classes at once:
//   dA1 = (dU W2^T) . [H1 > 0]   (rows x N, never materialised)
//   gW1[:, block g] += X[rows of g]^T dA1[rows of g]      gb1[block g] += column sums of dA1[rows of g]
// X rows x M (M = 128 or 256), H1 rows x N (N = 256), dU rows x Cw (Cw % 4 == 0, <= 64), W2 N x Cw.
int gs_mlp_bwd_grouped_f32(int32_t G, const int32_t* seg, const int32_t* out_block, int32_t M, int32_t N, int32_t Cw,
                           int32_t total_rows, const float* X, int64_t ldx, const float* H1, int64_t ldh,
                           const float* dU, int64_t ldu, const float* W2, int64_t ldw2, float* gW1, int64_t ldc,
                           float* gb1, int precision, void* stream) {
  GS_REQUIRE(G > 0 && seg && out_block && X && H1 && dU && W2 && gW1 && gb1 && total_rows > 0);
  GS_REQUIRE((M == 128 || M == 256) && N == 256 && Cw >= 4 && Cw <= gt::kMaxCw && Cw % 4 == 0 &&
             (precision == 1 || precision == 2));
  GS_REQUIRE(ldx >= M && ldh >= N && ldu >= Cw && ldw2 >= Cw && ldc >= N && ldx % 4 == 0 && ldh % 4 == 0 && ldu % 4 == 0 &&
             ldc % 4 == 0);
  GS_REQUIRE(gt_aligned16(X) && gt_aligned16(H1) && gt_aligned16(dU) && gt_aligned16(gW1) && gt_aligned16(gb1));
  CUtensorMap ma, mb;
  int rc = gt::encode_2d(&ma, X, total_rows, M, ldx, 32);
  if (rc) return rc;
  rc = gt::encode_2d(&mb, dU, total_rows, Cw, ldu, Cw);
  if (rc) return rc;
  gt::Params p{total_rows, (total_rows + gt::KR - 1) / gt::KR, G, seg, out_block, M, N, gW1, ldc, H1, ldh, W2, ldw2, Cw, gb1};
  return gt::dispatch(M / 64, 2, 4, precision, ma, mb, p, gs::as_stream(stream));
}

}  // extern "C"
