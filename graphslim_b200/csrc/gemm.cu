// Dense fp32 contractions, SIMT path (exact fp32 FMA).
//
//   C = alpha * op(A) * op(B) + beta * C          (row-major, explicit leading dimensions)
//
// These replace the ATen matmuls of the condense model and PGE (graphslim/models/sgc.py:39,49,
// models/layers.py:40-46,377, models/parametrized_adj.py:57-71) and every product autograd
// derives from them.  The large synthetic-side products are routed to the tcgen05 kernels in
// gemm_tc.cu when `precision` asks for it; this file is the exact-fp32 path and handles all the
// small / oddly shaped products (bias outer products, class-segmented weight gradients).
#include "common.cuh"

namespace gs {

int gemm_tc_dispatch(int ta, int tb, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float beta, float* C, int64_t ldc, int precision, void* workspace,
                     int64_t workspace_bytes, const float* bias, int relu, const float* mask, int64_t ldmask,
                     cudaStream_t st);
int64_t gemm_tc_workspace_bytes(int M, int N, int K, int precision);
int gemm_tc_grouped_dispatch(int G, const int32_t* seg, const int32_t* out_block, int M, int N, int K_total,
                             const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                             int precision, void* workspace, int64_t workspace_bytes, cudaStream_t st);

struct GemmProblem {
  int ta, tb;
  int M, N, K;
  float alpha, beta;
  const float* A;
  int64_t lda;
  const float* B;
  int64_t ldb;
  float* C;
  int64_t ldc;
  // split-K (groups == nullptr): blockIdx.z owns K range [z*kchunk, min(K,(z+1)*kchunk)), atomics when splits > 1
  int splits, kchunk;
  // grouped TN (seg != nullptr): blockIdx.z = group g; rows seg[g]..seg[g+1] are the K range,
  // output column block out_block[g]
  const int32_t* seg;
  const int32_t* out_block;
  // fused epilogue (splits == 1, not grouped): v += bias[col]; relu; v = mask[row,col] > 0 ? v : 0
  const float* bias;
  int relu;
  const float* mask;
  int64_t ldmask;
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(GemmProblem p) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  const float* A = p.A;
  const float* B = p.B;
  float* C = p.C;
  int kbeg = 0, kend = p.K;
  bool atomic = false;
  if (p.seg) {
    const int g = blockIdx.z;
    const int r0 = p.seg[g], r1 = p.seg[g + 1];
    A += (int64_t)r0 * p.lda;  // ta == 1: A is (K x M)
    B += (int64_t)r0 * p.ldb;  // tb == 0: B is (K x N)
    C += (int64_t)p.out_block[g] * p.N;
    kend = r1 - r0;
  } else if (p.splits > 1) {
    kbeg = blockIdx.z * p.kchunk;
    kend = min(p.K, kbeg + p.kchunk);
    atomic = true;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile (BM x BK) as As[k][m]
    if (p.ta) {
      for (int i = tid; i < BM * BK; i += NT) {
        const int m = i % BM, k = i / BM;
        const int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < p.M && gk < kend) ? __ldg(A + (int64_t)gk * p.lda + gm) : 0.f;
      }
    } else {
      for (int i = tid; i < BM * BK; i += NT) {
        const int k = i % BK, m = i / BK;
        const int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < p.M && gk < kend) ? __ldg(A + (int64_t)gm * p.lda + gk) : 0.f;
      }
    }
    // ---- stage B tile (BK x BN) as Bs[k][n]
    if (p.tb) {
      for (int i = tid; i < BN * BK; i += NT) {
        const int k = i % BK, n = i / BK;
        const int gn = n0 + n, gk = k0 + k;
        Bs[k][n] = (gn < p.N && gk < kend) ? __ldg(B + (int64_t)gn * p.ldb + gk) : 0.f;
      }
    } else {
      for (int i = tid; i < BN * BK; i += NT) {
        const int n = i % BN, k = i / BN;
        const int gn = n0 + n, gk = k0 + k;
        Bs[k][n] = (gn < p.N && gk < kend) ? __ldg(B + (int64_t)gk * p.ldb + gn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= p.N) continue;
      float* c = C + (int64_t)gm * p.ldc + gn;
      const float v = p.alpha * acc[i][j];
      if (atomic) {
        atomicAdd(c, v);
      } else {
        float o = (p.beta == 0.f) ? v : fmaf(p.beta, *c, v);
        if (p.bias) o += __ldg(p.bias + gn);
        if (p.relu) o = fmaxf(o, 0.f);
        if (p.mask) o = __ldg(p.mask + (int64_t)gm * p.ldmask + gn) > 0.f ? o : 0.f;
        *c = o;
      }
    }
  }
}

__global__ void scale_matrix_kernel(int M, int N, float* C, int64_t ldc, float beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  float* c = C + (i / N) * ldc + (i % N);
  *c = (beta == 0.f) ? 0.f : *c * beta;
}

// ------------------------------------------------------------------------------------------- skinny products
// One dimension of the product is at most 16: the class dimension of the 7-class graphs (Cora, Flickr), bias outer
// products with K = 1.  The tiled kernel above spends such a product on a single column of CTAs with a serial K loop
// (~100 us for the 446 x 446 x 7 propagation of the Flickr-shape inner loop, profiles/r2_launches_flickr.csv); the
// three kernels below cover them at memory speed.  Exact fp32 FMA, no atomics (bit-reproducible run to run).
constexpr int kSkinny = 16;

__device__ __forceinline__ float skinny_epilogue(const GemmProblem& p, float acc, int gm, int gn, float* c) {
  const float v = p.alpha * acc;
  float o = (p.beta == 0.f) ? v : fmaf(p.beta, *c, v);
  if (p.bias) o += __ldg(p.bias + gn);
  if (p.relu) o = fmaxf(o, 0.f);
  if (p.mask) o = __ldg(p.mask + (int64_t)gm * p.ldmask + gn) > 0.f ? o : 0.f;
  return o;
}

// (a) N <= NMAX, A not transposed: a warp owns RW consecutive rows, its lanes stride K (coalesced), op(B) sits in shared
// memory as Bs[n][k] so that consecutive lanes read consecutive words.  VEC: float4 loads of A (lda % 4 == 0, A aligned).
template <int NMAX, bool VEC>
__global__ void __launch_bounds__(256) gemm_skinny_n_kernel(GemmProblem p) {
  constexpr int KC = 512, RW = 4, STEP = VEC ? 128 : 32;
  __shared__ __align__(16) float Bs[NMAX][KC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrb = (p.M + 8 * RW - 1) / (8 * RW);
  const int nchunks = (p.K + KC - 1) / KC;
  for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
    const int m0 = (rb * 8 + warp) * RW;
    float acc[RW][NMAX];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int n = 0; n < NMAX; ++n) acc[r][n] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int kc = c * KC, kn = min(KC, p.K - kc);
      const int kpad = (kn + STEP - 1) / STEP * STEP;
      if (nchunks > 1 || rb == (int)blockIdx.x) {          // a single chunk stays resident across row blocks
        __syncthreads();
        for (int i = threadIdx.x; i < NMAX * kpad; i += 256) {
          const int n = i / kpad, k = i - n * kpad;
          float v = 0.f;
          if (n < p.N && k < kn) v = p.tb ? __ldg(p.B + (int64_t)n * p.ldb + kc + k) : __ldg(p.B + (int64_t)(kc + k) * p.ldb + n);
          Bs[n][k] = v;
        }
        __syncthreads();
      }
      for (int k = (VEC ? lane * 4 : lane); k < kn; k += STEP) {
        if constexpr (VEC) {
          float4 a[RW];
#pragma unroll
          for (int r = 0; r < RW; ++r) {
            a[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < p.M) {
              a[r] = __ldg(reinterpret_cast<const float4*>(p.A + (int64_t)(m0 + r) * p.lda + kc + k));
              if (k + 3 >= kn) {                           // the tail of the last vector lies beyond K (inside lda)
                if (k + 1 >= kn) a[r].y = 0.f;
                if (k + 2 >= kn) a[r].z = 0.f;
                a[r].w = 0.f;
              }
            }
          }
#pragma unroll
          for (int n = 0; n < NMAX; ++n) {
            if (n < p.N) {
              const float4 b = *reinterpret_cast<const float4*>(&Bs[n][k]);
#pragma unroll
              for (int r = 0; r < RW; ++r)
                acc[r][n] = fmaf(a[r].x, b.x, fmaf(a[r].y, b.y, fmaf(a[r].z, b.z, fmaf(a[r].w, b.w, acc[r][n]))));
            }
          }
        } else {
          float a[RW];
#pragma unroll
          for (int r = 0; r < RW; ++r) a[r] = (m0 + r < p.M) ? __ldg(p.A + (int64_t)(m0 + r) * p.lda + kc + k) : 0.f;
#pragma unroll
          for (int n = 0; n < NMAX; ++n) {
            if (n < p.N) {
              const float b = Bs[n][k];
#pragma unroll
              for (int r = 0; r < RW; ++r) acc[r][n] = fmaf(a[r], b, acc[r][n]);
            }
          }
        }
      }
    }
    // warp totals; lane (r * NMAX + n) % 32 keeps output (r, n)
    float mine[RW * NMAX / 32 > 0 ? RW * NMAX / 32 : 1];
#pragma unroll
    for (int i = 0; i < RW * NMAX / 32; ++i) mine[i] = 0.f;
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int n = 0; n < NMAX; ++n) {
        float v = acc[r][n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (((r * NMAX + n) & 31) == lane) mine[(r * NMAX + n) >> 5] = v;
      }
#pragma unroll
    for (int i = 0; i < RW * NMAX / 32; ++i) {
      const int o = i * 32 + lane, r = o / NMAX, n = o % NMAX;
      const int gm = m0 + r;
      if (gm < p.M && n < p.N) {
        float* c = p.C + (int64_t)gm * p.ldc + n;
        *c = skinny_epilogue(p, mine[i], gm, n, c);
      }
    }
  }
}

// (b) N <= NMAX, A transposed (K x M in memory), B (K x N): lanes own 32 consecutive columns m of A (coalesced), the 8
// warps of the CTA interleave the K range and their partial sums are added in warp order through shared memory.
// Grouped form (p.seg): blockIdx.z = class g, K range = rows seg[g]..seg[g+1], output column block out_block[g].
template <int NMAX>
__global__ void __launch_bounds__(256) gemm_skinny_tn_kernel(GemmProblem p) {
  __shared__ float red[8][NMAX][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 32 + lane;
  const float* A = p.A;
  const float* B = p.B;
  float* C = p.C;
  int kend = p.K;
  if (p.seg) {
    const int g = blockIdx.z;
    const int r0 = p.seg[g], r1 = p.seg[g + 1];
    A += (int64_t)r0 * p.lda;
    B += (int64_t)r0 * p.ldb;
    C += (int64_t)p.out_block[g] * p.N;
    kend = r1 - r0;
  }
  float acc[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
  const bool live = m < p.M;
  int k = warp;
  for (; k + 24 < kend; k += 32) {                          // four rows in flight per warp
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = live ? __ldg(A + (int64_t)(k + 8 * u) * p.lda + m) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < p.N) acc[n] = fmaf(a[u], __ldg(B + (int64_t)(k + 8 * u) * p.ldb + n), acc[n]);
  }
  for (; k < kend; k += 8) {
    const float a = live ? __ldg(A + (int64_t)k * p.lda + m) : 0.f;
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < p.N) acc[n] = fmaf(a, __ldg(B + (int64_t)k * p.ldb + n), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < NMAX; ++n) red[warp][n][lane] = acc[n];
  __syncthreads();
  for (int n = warp; n < p.N; n += 8) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][n][lane];
    if (live) {
      float* c = C + (int64_t)m * p.ldc + n;
      *c = skinny_epilogue(p, v, m, n, c);
    }
  }
}

// (c) K <= 16: every output is a short dot product.  threadIdx.x -> column n (b[k] in registers), threadIdx.y strides the
// rows of a 64-row chunk whose A values are staged in shared memory (read back as broadcasts).
__global__ void __launch_bounds__(256) gemm_skinny_k_kernel(GemmProblem p) {
  constexpr int RC = 64;
  __shared__ __align__(16) float As[RC][kSkinny];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
  float b[kSkinny];
#pragma unroll
  for (int k = 0; k < kSkinny; ++k) {
    b[k] = 0.f;
    if (k < p.K && n < p.N) b[k] = p.tb ? __ldg(p.B + (int64_t)n * p.ldb + k) : __ldg(p.B + (int64_t)k * p.ldb + n);
  }
  const int nchunks = (p.M + RC - 1) / RC;
  for (int ch = blockIdx.y; ch < nchunks; ch += gridDim.y) {
    const int m0 = ch * RC;
    __syncthreads();
    for (int i = tid; i < RC * kSkinny; i += nthreads) {
      const int r = i / kSkinny, k = i % kSkinny;
      float v = 0.f;
      if (m0 + r < p.M && k < p.K) v = p.ta ? __ldg(p.A + (int64_t)k * p.lda + m0 + r) : __ldg(p.A + (int64_t)(m0 + r) * p.lda + k);
      As[r][k] = v;
    }
    __syncthreads();
    if (n < p.N) {
      for (int r = threadIdx.y; r < RC && m0 + r < p.M; r += blockDim.y) {
        float v = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < kSkinny; k4 += 4) {
          if (k4 < p.K) {
            const float4 a = *reinterpret_cast<const float4*>(&As[r][k4]);
            v = fmaf(a.x, b[k4], fmaf(a.y, b[k4 + 1], fmaf(a.z, b[k4 + 2], fmaf(a.w, b[k4 + 3], v))));
          }
        }
        float* c = p.C + (int64_t)(m0 + r) * p.ldc + n;
        *c = skinny_epilogue(p, v, m0 + r, n, c);
      }
    }
  }
}

// Routes a product to one of the skinny kernels; GS_ENOSYS when none applies.
static int launch_skinny(GemmProblem& p, int groups, cudaStream_t st) {
  if (p.splits > 1) return GS_ENOSYS;
  if (p.seg) {
    if (p.N > kSkinny) return GS_ENOSYS;
    dim3 grid((p.M + 31) / 32, 1, groups);
    if (p.N <= 8) gemm_skinny_tn_kernel<8><<<grid, 256, 0, st>>>(p);
    else gemm_skinny_tn_kernel<16><<<grid, 256, 0, st>>>(p);
    return finish_launch("gemm_skinny_tn");
  }
  if (p.N <= kSkinny && p.K > kSkinny) {
    if (p.ta) {
      if (p.tb) return GS_ENOSYS;
      dim3 grid((p.M + 31) / 32, 1, 1);
      if (p.N <= 8) gemm_skinny_tn_kernel<8><<<grid, 256, 0, st>>>(p);
      else gemm_skinny_tn_kernel<16><<<grid, 256, 0, st>>>(p);
      return finish_launch("gemm_skinny_tn");
    }
    const bool vec = (p.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(p.A) & 15) == 0;
    const int nrb = (p.M + 31) / 32;
    const int grid = std::min(nrb, 8 * kNumSMs);
    if (p.N <= 8) {
      if (vec) gemm_skinny_n_kernel<8, true><<<grid, 256, 0, st>>>(p);
      else gemm_skinny_n_kernel<8, false><<<grid, 256, 0, st>>>(p);
    } else {
      if (vec) gemm_skinny_n_kernel<16, true><<<grid, 256, 0, st>>>(p);
      else gemm_skinny_n_kernel<16, false><<<grid, 256, 0, st>>>(p);
    }
    return finish_launch("gemm_skinny_n");
  }
  if (p.K <= kSkinny && p.K > 0) {
    const int bx = p.N >= 256 ? 256 : std::max(32, (p.N + 31) / 32 * 32);
    const int by = std::max(1, 256 / bx);
    dim3 block(bx, by, 1);
    const int nchunks = (p.M + 63) / 64;
    dim3 grid((p.N + bx - 1) / bx, std::min(nchunks, 4 * kNumSMs), 1);
    gemm_skinny_k_kernel<<<grid, block, 0, st>>>(p);
    return finish_launch("gemm_skinny_k");
  }
  return GS_ENOSYS;
}

static int launch_simt(GemmProblem& p, int gz, cudaStream_t st) {
  const bool big = (int64_t)p.M * p.N >= (int64_t)512 * 512 && p.M >= 128 && p.N >= 128;
  if (big) {
    dim3 grid((p.N + 127) / 128, (p.M + 127) / 128, gz);
    gemm_simt_kernel<128, 128, 8, 8, 8><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((p.N + 63) / 64, (p.M + 63) / 64, gz);
    gemm_simt_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, st>>>(p);
  }
  return finish_launch("gemm_simt");
}

}  // namespace gs

extern "C" {

int gs_gemm_epi_f32(int ta, int tb, int32_t M, int32_t N, int32_t K, float alpha, const float* A, int64_t lda,
                    const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int relu,
                    const float* mask, int64_t ldmask, int precision, void* workspace, int64_t workspace_bytes,
                    void* stream) {
  GS_REQUIRE(M >= 0 && N >= 0 && K >= 0 && C && ldc >= N);
  GS_REQUIRE(K == 0 || (A && B));
  GS_REQUIRE(K == 0 || lda >= (ta ? M : K));
  GS_REQUIRE(K == 0 || ldb >= (tb ? K : N));
  GS_REQUIRE(!mask || ldmask >= N);
  if (M == 0 || N == 0) return GS_OK;
  const bool epi = bias || relu || mask;
  cudaStream_t st = gs::as_stream(stream);
  if (precision != 0) {
    const int rc = gs::gemm_tc_dispatch(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, precision, workspace,
                                        workspace_bytes, bias, relu, mask, ldmask, st);
    if (rc != GS_ENOSYS) return rc;  // GS_ENOSYS: shape not covered by the tensor-core kernels
  }
  gs::GemmProblem p{ta, tb, M, N, K, alpha, beta, A, lda, B, ldb, C, ldc, 1, K, nullptr, nullptr, bias, relu, mask, ldmask};
  if (N <= gs::kSkinny || (K <= gs::kSkinny && K > 0)) {   // one dimension <= 16: dedicated memory-speed kernels
    const int rc = gs::launch_skinny(p, 1, st);
    if (rc != GS_ENOSYS) return rc;
  }
  // split K when the output alone cannot fill the machine (e.g. dW = dY^T H with K = N'^2); a fused epilogue needs
  // the complete sum in one thread, so those products stay unsplit
  const int64_t tiles = (int64_t)((M + 63) / 64) * ((N + 63) / 64);
  int splits = 1;
  // (K >= 512: the 70 x 1433 x 7 logits product of the Cora-shape inner loop ran 238 us as one serial K loop on two
  // CTAs -- 70 % of an inner step; profiles/r1_gemm_shapes_cora.json)
  if (K >= 512 && tiles < 2 * gs::kNumSMs && !epi) {
    splits = (int)((2 * gs::kNumSMs + tiles - 1) / tiles);
    const int max_splits = K >= 2048 ? (K + 511) / 512 : (K + 127) / 128;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  if (splits > 1) {
    int kchunk = (K + splits - 1) / splits;
    kchunk = (kchunk + 15) / 16 * 16;
    splits = (K + kchunk - 1) / kchunk;
    p.splits = splits;
    p.kchunk = kchunk;
    if (splits > 1) {
      const int64_t n = (int64_t)M * N;
      gs::scale_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(M, N, C, ldc, beta);
      const int rc = gs::finish_launch("scale_matrix");
      if (rc) return rc;
    }
  }
  return gs::launch_simt(p, p.splits, st);
}

int gs_gemm_f32(int ta, int tb, int32_t M, int32_t N, int32_t K, float alpha, const float* A, int64_t lda,
                const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int precision, void* workspace,
                int64_t workspace_bytes, void* stream) {
  return gs_gemm_epi_f32(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, nullptr, 0, nullptr, 0, precision,
                         workspace, workspace_bytes, stream);
}

int64_t gs_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K, int precision) {
  return gs::gemm_tc_workspace_bytes(M, N, K, precision);
}

int gs_gemm_grouped_tn_f32(int32_t G, const int32_t* seg, const int32_t* out_block, int32_t M, int32_t N,
                           int32_t K_total, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                           int64_t ldc, int precision, void* workspace, int64_t workspace_bytes, void* stream) {
  GS_REQUIRE(G >= 0 && seg && out_block && M >= 0 && N >= 0 && K_total >= 0 && A && B && C && lda >= M && ldb >= N);
  if (G == 0 || M == 0 || N == 0) return GS_OK;
  if (precision != 0) {
    const int rc = gs::gemm_tc_grouped_dispatch(G, seg, out_block, M, N, K_total, A, lda, B, ldb, C, ldc, precision,
                                                workspace, workspace_bytes, gs::as_stream(stream));
    if (rc != GS_ENOSYS) return rc;
  }
  gs::GemmProblem p{1, 0, M, N, 0, 1.f, 0.f, A, lda, B, ldb, C, ldc, 1, 0, seg, out_block, nullptr, 0, nullptr, 0};
  if (N <= gs::kSkinny) {
    const int rc = gs::launch_skinny(p, G, gs::as_stream(stream));
    if (rc != GS_ENOSYS) return rc;
  }
  return gs::launch_simt(p, G, gs::as_stream(stream));
}

}  // extern "C"
