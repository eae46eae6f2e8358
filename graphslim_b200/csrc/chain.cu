// One persistent cooperative kernel that runs a recorded list of small dense operations back to back.
//
// The condense-model training steps of the inner loop (graphslim/condensation/gcond.py:63-72: forward on the
// synthetic graph, nll gradients, one Adam step) are ~35 launches of kernels that each keep a dozen CTAs busy for
// 3-20 us: every product has at most N' = |synthetic nodes| rows (909 at the ogbn-arxiv shape, 70 at Cora).  Even
// replayed from a CUDA graph the step is a chain of dependent launches, 180 us at the arxiv shape and 65 % of a
// Cora-shape epoch (profiles/r2_launches_*.csv).  Here the host records the step once as an array of descriptors
// (graphslim_b200/chain.py) and this kernel walks the array with every SM taking part in every operation and a grid
// barrier where an operation depends on the previous ones.
//
// Operands written by one operation are read by later ones inside the same launch, so every load here is a plain
// (coherent after grid.sync()) load: no __ldg / const __restrict__ non-coherent path.
//
// Arithmetic: exact fp32 FMA at every gemm_precision; every reduction runs in a fixed order (no atomics), so a
// program gives the same bits on every run.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace gs {
namespace chain {

constexpr int kThreads = 256, kWarps = 8;
constexpr int TS = 32;             // output tile edge; also the k-chunk a warp stages at a time
constexpr int LDS_ = 36;           // shared row stride in floats (16-byte aligned rows for the broadcast float4 reads)

enum Kind { GEMM = 0, SOFTMAX_RESIDUAL = 1, COLSUM = 2, ADAM_TABLE = 3, COUNTER_ADD = 4, FILL = 5 };

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// C = epi(alpha op(A) op(B) + beta C).  A CTA owns a 32 x 32 output tile; its 8 warps interleave the 32-deep k-chunks
// (lane = output column, 32 row accumulators per lane, op(A) chunk broadcast from the warp's shared tile), then the
// warps' partial tiles are added in warp order.
__device__ void run_gemm(const gs_chain_op& op, float* sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* A = reinterpret_cast<const float*>(op.A);
  const float* B = reinterpret_cast<const float*>(op.B);
  float* C = reinterpret_cast<float*>(op.C);
  const float* bias = reinterpret_cast<const float*>(op.bias);
  const float* mask = reinterpret_cast<const float*>(op.mask);
  const int M = op.M, N = op.N, K = op.K;
  const int tn = (N + TS - 1) / TS, tiles = ((M + TS - 1) / TS) * tn;
  float* As = sm + warp * (TS * LDS_);                     // this warp's [kk][m] tile
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int m0 = (t / tn) * TS, n0 = (t % tn) * TS;
    const int gn = n0 + lane;
    float acc[TS];
#pragma unroll
    for (int m = 0; m < TS; ++m) acc[m] = 0.f;
    for (int k0 = warp * TS; k0 < K; k0 += kWarps * TS) {
      float b[TS];
      if (op.tb) {
        const float* bp = B + (int64_t)(gn < N ? gn : 0) * op.ldb + k0;
#pragma unroll
        for (int kk = 0; kk < TS; ++kk) b[kk] = (gn < N && k0 + kk < K) ? *(bp + kk) : 0.f;
      } else {
#pragma unroll
        for (int kk = 0; kk < TS; ++kk) b[kk] = (gn < N && k0 + kk < K) ? *(B + (int64_t)(k0 + kk) * op.ldb + gn) : 0.f;
      }
      if (op.ta) {                                          // A is K x M: a row of the chunk is contiguous in m
#pragma unroll 8
        for (int kk = 0; kk < TS; ++kk)
          As[kk * LDS_ + lane] = (k0 + kk < K && m0 + lane < M) ? *(A + (int64_t)(k0 + kk) * op.lda + m0 + lane) : 0.f;
      } else {                                              // A is M x K: lanes run along k
#pragma unroll 8
        for (int mm = 0; mm < TS; ++mm)
          As[lane * LDS_ + mm] = (m0 + mm < M && k0 + lane < K) ? *(A + (int64_t)(m0 + mm) * op.lda + k0 + lane) : 0.f;
      }
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < TS; ++kk) {
        const float bk = b[kk];
#pragma unroll
        for (int j = 0; j < TS / 4; ++j) {
          const float4 a = *reinterpret_cast<const float4*>(As + kk * LDS_ + 4 * j);
          acc[4 * j + 0] = fmaf(a.x, bk, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(a.y, bk, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(a.z, bk, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(a.w, bk, acc[4 * j + 3]);
        }
      }
      __syncwarp();
    }
    __syncthreads();                                        // every warp is done with its staging tile
    float* red = sm;                                        // [warp][m][lane], 32 KB of the 36 KB staging area
#pragma unroll
    for (int m = 0; m < TS; ++m) red[(warp * TS + m) * TS + lane] = acc[m];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TS / kWarps; ++i) {
      const int m = warp + kWarps * i, gm = m0 + m;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += red[(w * TS + m) * TS + lane];
      if (gm < M && gn < N) {
        float* c = C + (int64_t)gm * op.ldc + gn;
        float o = op.alpha * v;
        if (op.beta != 0.f) o = fmaf(op.beta, *c, o);
        if (bias) o += *(bias + gn);
        if (op.relu) o = fmaxf(o, 0.f);
        if (mask) o = *(mask + (int64_t)gm * op.ldmask + gn) > 0.f ? o : 0.f;
        *c = o;
      }
    }
    __syncthreads();
  }
}

// S = softmax(Z) by rows, R = (S - onehot(label)) * row_scale   (one warp per row; same arithmetic as
// softmax_residual_kernel in elementwise.cu)
__device__ void run_softmax_residual(const gs_chain_op& op) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* Z = reinterpret_cast<const float*>(op.A);
  const int32_t* label = reinterpret_cast<const int32_t*>(op.B);
  const float* row_scale = reinterpret_cast<const float*>(op.bias);
  float* S = reinterpret_cast<float*>(op.C);
  float* R = reinterpret_cast<float*>(op.p5);
  const int rows = op.M, Cc = op.N;
  for (int row = blockIdx.x * kWarps + warp; row < rows; row += gridDim.x * kWarps) {
    const float* z = Z + (int64_t)row * op.lda;
    float mx = -INFINITY;
    for (int c = lane; c < Cc; c += 32) mx = fmaxf(mx, z[c]);
    mx = warp_max_f(mx);
    float sum = 0.f;
    for (int c = lane; c < Cc; c += 32) sum += expf(z[c] - mx);
    sum = warp_sum_f(sum);
    const float lse = mx + logf(sum);
    const int y = label[row];
    const float sc = row_scale ? row_scale[row] : 1.f;
    for (int c = lane; c < Cc; c += 32) {
      const float s = expf(z[c] - lse);
      if (S) S[(int64_t)row * Cc + c] = s;
      if (R) R[(int64_t)row * Cc + c] = (s - (c == y ? 1.f : 0.f)) * sc;
    }
  }
}

// out[c] = sum over rows of X[r][c]: a CTA owns 32 columns, its warps interleave the rows, partials added in warp order
__device__ void run_colsum(const gs_chain_op& op, float* sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* X = reinterpret_cast<const float*>(op.A);
  float* out = reinterpret_cast<float*>(op.C);
  const int rows = op.M, cols = op.N;
  const int blocks = (cols + 31) / 32;
  for (int t = blockIdx.x; t < blocks; t += gridDim.x) {
    const int c = t * 32 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < cols) {
      int r = warp;
      for (; r + 3 * kWarps < rows; r += 4 * kWarps) {
        s0 += X[(int64_t)r * op.lda + c];
        s1 += X[(int64_t)(r + kWarps) * op.lda + c];
        s2 += X[(int64_t)(r + 2 * kWarps) * op.lda + c];
        s3 += X[(int64_t)(r + 3 * kWarps) * op.lda + c];
      }
      for (; r < rows; r += kWarps) s0 += X[(int64_t)r * op.lda + c];
    }
    __syncthreads();
    sm[warp * 32 + lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0 && c < cols) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += sm[w * 32 + lane];
      out[c] = v;
    }
  }
}

// Adam with the step-dependent scalars read from a device table (same arithmetic as adam_table_kernel)
__device__ void run_adam_table(const gs_chain_op& op) {
  float* p = reinterpret_cast<float*>(op.C);
  const float* g = reinterpret_cast<const float*>(op.A);
  float* m = reinterpret_cast<float*>(op.p5);
  float* v = reinterpret_cast<float*>(op.p6);
  const float* table = reinterpret_cast<const float*>(op.B);
  const int t = *reinterpret_cast<const int32_t*>(op.p7);
  const float step_size = table[2 * t], bc2_sqrt = table[2 * t + 1];
  const float om_beta1 = op.alpha, beta2 = op.beta, om_beta2 = op.f0, eps = op.f1;
  const int64_t n = op.lda;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float gi = g[i];
    const float mi = m[i] + om_beta1 * (gi - m[i]);
    const float vi = v[i] * beta2 + (om_beta2 * gi) * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

__global__ void __launch_bounds__(kThreads) chain_kernel(const gs_chain_op* __restrict__ ops, int n_ops) {
  cg::grid_group grid = cg::this_grid();
  __shared__ __align__(16) float sm[kWarps * TS * LDS_];
  __shared__ gs_chain_op sop;
  for (int o = 0; o < n_ops; ++o) {
    if (o > 0 && ops[o].sync_before) grid.sync();
    __syncthreads();                                        // previous operation done with `sm` / `sop` in this CTA
    if (threadIdx.x < (int)(sizeof(gs_chain_op) / 4))
      reinterpret_cast<uint32_t*>(&sop)[threadIdx.x] = reinterpret_cast<const uint32_t*>(ops + o)[threadIdx.x];
    __syncthreads();
    switch (sop.kind) {
      case GEMM: run_gemm(sop, sm); break;
      case SOFTMAX_RESIDUAL: run_softmax_residual(sop); break;
      case COLSUM: run_colsum(sop, sm); break;
      case ADAM_TABLE: run_adam_table(sop); break;
      case COUNTER_ADD:
        if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<int32_t*>(sop.C) += sop.M;
        break;
      case FILL: {
        float* c = reinterpret_cast<float*>(sop.C);
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < sop.lda; i += (int64_t)gridDim.x * kThreads)
          c[i] = sop.alpha;
        break;
      }
      default: break;
    }
  }
}

}  // namespace chain
}  // namespace gs

extern "C" {

int gs_chain_run_f32(const gs_chain_op* ops_dev, int32_t n_ops, void* stream) {
  GS_REQUIRE(ops_dev && n_ops > 0);
  static int grid = 0;
  if (grid == 0) {
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) {
      gs::set_error_msg("gs_chain_run_f32: device does not support cooperative launches");
      return GS_ENOSYS;
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs::chain::chain_kernel, gs::chain::kThreads, 0);
    if (e != cudaSuccess || per_sm < 1) {
      gs::set_error("gs_chain_run_f32: occupancy query", e);
      return (int)(e != cudaSuccess ? e : cudaErrorUnknown);
    }
    grid = sms * (per_sm < 2 ? per_sm : 2);
  }
  const gs_chain_op* ops = ops_dev;
  int n = n_ops;
  void* args[] = {(void*)&ops, (void*)&n};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)gs::chain::chain_kernel, dim3(grid), dim3(gs::chain::kThreads), args,
                                              0, gs::as_stream(stream));
  if (e != cudaSuccess) {
    gs::set_error("gs_chain_run_f32: cooperative launch", e);
    return (int)e;
  }
  return gs::finish_launch("chain");
}

}  // extern "C"
