// Small fused kernels of the condense model, the gradient-matching loss (K8), the dense GCN
// normalisation (K6) and the optimiser.  All bandwidth-trivial; the point of writing them by hand
// is to keep the whole outer step a fixed sequence of our own launches on one stream.
#include "common.cuh"

namespace gs {

// ---------------------------------------------------------------- bias / relu
__global__ void bias_act_kernel(int rows, int cols, float* __restrict__ Z, int64_t ldz, const float* __restrict__ bias,
                                int relu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  float v = Z[(int64_t)r * ldz + c];
  if (bias) v += __ldg(bias + c);
  if (relu) v = fmaxf(v, 0.f);
  Z[(int64_t)r * ldz + c] = v;
}

__global__ void relu_mask_kernel(int rows, int groups, int cols, float* __restrict__ D, const float* __restrict__ H,
                                 int64_t ldh) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)rows * groups * cols;
  if (i >= total) return;
  const int c = (int)(i % cols);
  const int r = (int)(i / ((int64_t)groups * cols));
  if (!(__ldg(H + (int64_t)r * ldh + c) > 0.f)) D[i] = 0.f;
}

// ---------------------------------------------------------------- softmax family (one warp per row)
__global__ void softmax_residual_kernel(int rows, int C, const float* __restrict__ Z, int64_t ldz,
                                        const int32_t* __restrict__ label, const float* __restrict__ row_scale,
                                        float* __restrict__ S, float* __restrict__ R, float* __restrict__ nll) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* z = Z + (int64_t)row * ldz;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(z[c] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  const int y = label[row];
  const float sc = row_scale ? row_scale[row] : 1.f;
  for (int c = lane; c < C; c += 32) {
    const float s = expf(z[c] - lse);
    if (S) S[(int64_t)row * C + c] = s;
    if (R) R[(int64_t)row * C + c] = (s - (c == y ? 1.f : 0.f)) * sc;
  }
  if (nll && lane == 0) nll[row] = lse - z[y];
}

__global__ void expand_class_blocks_kernel(int rows, int C, int nblk, const float* __restrict__ R,
                                           const int32_t* __restrict__ blk, float* __restrict__ E) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t W = (int64_t)nblk * C;
  if (i >= (int64_t)rows * W) return;
  const int r = (int)(i / W);
  const int w = (int)(i % W);
  const int b = w / C, c = w % C;
  E[i] = (b == blk[r]) ? R[(int64_t)r * C + c] : 0.f;
}

__global__ void pick_class_blocks_kernel(int rows, int C, int nblk, const float* __restrict__ Zf,
                                         const int32_t* __restrict__ blk, float* __restrict__ Q) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * C) return;
  const int r = (int)(i / C), c = (int)(i % C);
  const int b = blk[r];
  Q[i] = (b >= 0) ? Zf[(int64_t)r * nblk * C + (int64_t)b * C + c] : 0.f;
}

__global__ void softmax_jvp_kernel(int rows, int C, const float* __restrict__ S, const float* __restrict__ Q,
                                   const float* __restrict__ row_scale, float* __restrict__ dZ) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float sc = row_scale ? row_scale[row] : 1.f;
  const float* s = S + (int64_t)row * C;
  const float* q = Q + (int64_t)row * C;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot = fmaf(s[c], q[c] * sc, dot);
  dot = warp_sum(dot);
  for (int c = lane; c < C; c += 32) dZ[(int64_t)row * C + c] = s[c] * (q[c] * sc - dot);
}

// ---------------------------------------------------------------- gradient matching
// 32 columns per CTA: lanes read 32 consecutive columns of a row (coalesced), the 8 warps interleave the rows and their
// partial sums are added in warp order through shared memory (one thread per column with a serial row loop took 96 us
// on the 500 x 1792 first-layer gradient of the Flickr shape: 14 CTAs, 500 dependent loads each).
__global__ void __launch_bounds__(256) match_col_stats_kernel(int rows, int cols, const float* __restrict__ gs_,
                                                              const float* __restrict__ gr, int64_t ld,
                                                              float* __restrict__ stats, int64_t sld) {
  __shared__ float red[4][8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float dot = 0.f, ns = 0.f, nr = 0.f, sq = 0.f;
  if (j < cols) {
    for (int r = w; r < rows; r += 8) {
      const float a = gs_[(int64_t)r * ld + j], b = gr[(int64_t)r * ld + j];
      dot = fmaf(a, b, dot);
      ns = fmaf(a, a, ns);
      nr = fmaf(b, b, nr);
      const float d = a - b;
      sq = fmaf(d, d, sq);
    }
  }
  red[0][w][lane] = dot;
  red[1][w][lane] = ns;
  red[2][w][lane] = nr;
  red[3][w][lane] = sq;
  __syncthreads();
  if (w < 4 && j < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[w][i][lane];
    stats[(int64_t)w * sld + j] = v;
  }
}

// One block per class.  metric 0 'ours', 1 'mse', 2 'cos'  (graphslim/condensation/utils.py:12-106)
__global__ void match_finalize_kernel(int metric, int n_par, const int32_t* __restrict__ par_off,
                                      const int32_t* __restrict__ par_width, const int32_t* __restrict__ par_is_bias,
                                      const float* __restrict__ coeff, const float* __restrict__ stats, int64_t sld,
                                      float* __restrict__ alpha, float* __restrict__ beta,
                                      float* __restrict__ loss_per_class) {
  const int c = blockIdx.x;
  const float co = coeff[c];
  __shared__ float red[3][32];
  __shared__ float tot[3];
  float l_dot = 0.f, l_ns = 0.f, l_nr = 0.f, l_loss = 0.f;
  if (metric == 2) {
    for (int p = 0; p < n_par; ++p) {
      const int w = par_width[p], base = par_off[p] + c * w;
      for (int o = threadIdx.x; o < w; o += blockDim.x) {
        l_dot += stats[base + o];
        l_ns += stats[sld + base + o];
        l_nr += stats[2 * sld + base + o];
      }
    }
    l_dot = warp_sum(l_dot);
    l_ns = warp_sum(l_ns);
    l_nr = warp_sum(l_nr);
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = l_dot;
      red[1][threadIdx.x >> 5] = l_ns;
      red[2][threadIdx.x >> 5] = l_nr;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int nw = blockDim.x >> 5;
      float a = threadIdx.x < nw ? red[0][threadIdx.x] : 0.f;
      float b = threadIdx.x < nw ? red[1][threadIdx.x] : 0.f;
      float d = threadIdx.x < nw ? red[2][threadIdx.x] : 0.f;
      a = warp_sum(a);
      b = warp_sum(b);
      d = warp_sum(d);
      if (threadIdx.x == 0) {
        tot[0] = a;
        tot[1] = b;
        tot[2] = d;
      }
    }
    __syncthreads();
  }
  for (int p = 0; p < n_par; ++p) {
    const int w = par_width[p], base = par_off[p] + c * w;
    for (int o = threadIdx.x; o < w; o += blockDim.x) {
      const int j = base + o;
      float al = 0.f, be = 0.f;
      if (metric == 0) {
        if (!par_is_bias[p]) {
          const float dot = stats[j], ns = sqrtf(stats[sld + j]), nr = sqrtf(stats[2 * sld + j]);
          const float den = ns * nr + 1e-6f;
          l_loss += co * (1.f - dot / den);
          al = -co / den;
          be = (ns > 0.f) ? co * dot * nr / (ns * den * den) : 0.f;
        }
      } else if (metric == 1) {
        l_loss += co * stats[3 * sld + j];
        al = -2.f * co;
        be = 2.f * co;
      } else {
        const float ns = sqrtf(tot[1]), nr = sqrtf(tot[2]);
        const float den = ns * nr + 1e-6f;
        al = -co / den;
        be = (ns > 0.f) ? co * tot[0] * nr / (ns * den * den) : 0.f;
      }
      alpha[j] = al;
      beta[j] = be;
    }
  }
  if (metric == 2) {
    if (threadIdx.x == 0) {
      const float ns = sqrtf(tot[1]), nr = sqrtf(tot[2]);
      loss_per_class[c] = co * (1.f - tot[0] / (ns * nr + 1e-6f));
    }
    return;
  }
  __syncthreads();
  l_loss = warp_sum(l_loss);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = l_loss;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    float a = threadIdx.x < nw ? red[0][threadIdx.x] : 0.f;
    a = warp_sum(a);
    if (threadIdx.x == 0) loss_per_class[c] = a;
  }
}

__global__ void sum_small_kernel(int n, const float* __restrict__ x, float* __restrict__ out) {
  // single warp, fixed order -> deterministic
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += x[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *out += s;
}

__global__ void match_apply_kernel(int rows, int cols, const float* __restrict__ gs_, const float* __restrict__ gr,
                                   int64_t ld, const float* __restrict__ alpha, const float* __restrict__ beta,
                                   float* __restrict__ G) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), j = (int)(i % cols);
  G[(int64_t)r * ld + j] = alpha[j] * gr[(int64_t)r * ld + j] + beta[j] * gs_[(int64_t)r * ld + j];
}

// ---------------------------------------------------------------- per-class column sums (bias gradients)
// out[out_block[g]*cols + c] += sum_{r in seg[g]..seg[g+1]} X[r, c]   (out zeroed by the caller).
// The row range seg[0]..seg[G] is cut into 64-row runs, one per warp (blockIdx.y * 8 + warp), a lane per column of the
// block's 32-column tile; a run keeps eight row loads in flight, follows the segment boundaries it crosses and commits
// one atomic per (segment, column).  Parallelism comes from the rows, not from the number of groups.
constexpr int kColsumRun = 64;
__global__ void __launch_bounds__(256)
segment_colsum_kernel(int G, const int32_t* __restrict__ seg, const int32_t* __restrict__ out_block, int cols,
                      const float* __restrict__ X, int64_t ldx, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool live = c < cols;
  const int r_last = __ldg(seg + G);
  int r = __ldg(seg) + (blockIdx.y * 8 + w) * kColsumRun;
  const int r_end = min(r + kColsumRun, r_last);
  if (r >= r_end) return;
  int lo = 0, hi = G - 1;                 // last g with seg[g] <= r: the non-empty segment that holds row r
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(seg + mid) <= r) lo = mid;
    else hi = mid - 1;
  }
  int g = lo;
  const float* x = X + c;
  while (r < r_end) {
    const int s_end = min(r_end, __ldg(seg + g + 1));
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.f;
    if (live) {
      for (; r + 8 <= s_end; r += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(x + (int64_t)(r + i) * ldx);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += v[i];
      }
      for (; r < s_end; ++r) a[0] += __ldg(x + (int64_t)r * ldx);
      const float sum = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
      atomicAdd(out + (int64_t)__ldg(out_block + g) * cols + c, sum);
    }
    r = s_end;
    if (r < r_end) {
      ++g;
      while (__ldg(seg + g + 1) <= r) ++g;   // skip empty segments
    }
  }
}

// ---------------------------------------------------------------- dense GCN normalisation
// r_i = (1 + sum_j A_ij)^(-1/2);  Ahat_ij = (r_i * (A_ij + [i==j])) * r_j      (utils.py:429-439)
__global__ void dense_norm_rowsum_kernel(int n, const float* __restrict__ A, float* __restrict__ r) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += A[(int64_t)row * n + j];
  s = warp_sum(s) + 1.f;
  if (lane == 0) {
    const float v = 1.f / sqrtf(s);
    r[row] = isinf(v) ? 0.f : v;
  }
}
__global__ void dense_norm_scale_kernel(int n, const float* __restrict__ A, const float* __restrict__ r,
                                        float* __restrict__ Ahat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * n) return;
  const int a = (int)(i / n), b = (int)(i % n);
  const float m = A[i] + (a == b ? 1.f : 0.f);
  Ahat[i] = (r[a] * m) * r[b];
}
// work[0..n) = rowdot_i = sum_j dAhat_ij*Ahat_ij ; work[n..2n) = coldot_i = sum_j dAhat_ji*Ahat_ji
__global__ void dense_norm_bwd_dots_kernel(int n, const float* __restrict__ dAh, const float* __restrict__ Ah,
                                           float* __restrict__ work) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float rs = 0.f, cs = 0.f;
  for (int j = lane; j < n; j += 32) {
    rs = fmaf(dAh[(int64_t)row * n + j], Ah[(int64_t)row * n + j], rs);
    cs = fmaf(dAh[(int64_t)j * n + row], Ah[(int64_t)j * n + row], cs);
  }
  rs = warp_sum(rs);
  cs = warp_sum(cs);
  if (lane == 0) {
    work[row] = rs;
    work[n + row] = cs;
  }
}
__global__ void dense_norm_bwd_apply_kernel(int n, const float* __restrict__ dAh, const float* __restrict__ r,
                                            const float* __restrict__ work, float* __restrict__ dA) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * n) return;
  const int a = (int)(i / n), b = (int)(i % n);
  // dL/drho_a = -1/2 r_a^2 (rowdot_a + coldot_a)
  const float drho = -0.5f * r[a] * r[a] * (work[a] + work[n + a]);
  dA[i] = r[a] * r[b] * dAh[i] + drho;
}

// ---------------------------------------------------------------- optimiser
__global__ void adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float step_size, float bc2_sqrt, float om_beta1, float beta2,
                            float om_beta2, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  // torch _single_tensor_adam: exp_avg.lerp_(grad, 1-beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
  const float mi = m[i] + om_beta1 * (gi - m[i]);
  const float vi = v[i] * beta2 + (om_beta2 * gi) * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - step_size * (mi / denom);
}

// Adam step whose two step-dependent scalars come from device memory: table[2*t] = lr / (1 - beta1^(t+1)) and
// table[2*t+1] = sqrt(1 - beta2^(t+1)) for the zero-based step t = *step_dev.  Nothing in the launch depends on the
// step any more, so an inner-loop iteration can be captured once in a CUDA graph and replayed.
__global__ void adam_table_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, const float* __restrict__ table,
                                  const int32_t* __restrict__ step_dev, float om_beta1, float beta2, float om_beta2,
                                  float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = *step_dev;
  const float step_size = table[2 * t], bc2_sqrt = table[2 * t + 1];
  const float gi = g[i];
  const float mi = m[i] + om_beta1 * (gi - m[i]);
  const float vi = v[i] * beta2 + (om_beta2 * gi) * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - step_size * (mi / denom);
}

__global__ void counter_add_kernel(int32_t* c, int32_t inc) { *c += inc; }

__global__ void axpby_kernel(int64_t n, float a, const float* __restrict__ x, float b, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = (b == 0.f) ? a * x[i] : fmaf(a, x[i], b * y[i]);
}

static inline unsigned blocks_for(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace gs

extern "C" {
using namespace gs;

int gs_bias_act_f32(int32_t rows, int32_t cols, float* Z, int64_t ldz, const float* bias, int relu, void* stream) {
  GS_REQUIRE(rows >= 0 && cols >= 0 && Z && ldz >= cols);
  if ((int64_t)rows * cols == 0) return GS_OK;
  bias_act_kernel<<<blocks_for((int64_t)rows * cols), 256, 0, as_stream(stream)>>>(rows, cols, Z, ldz, bias, relu);
  return finish_launch("bias_act");
}

int gs_relu_mask_f32(int32_t rows, int32_t groups, int32_t cols, float* D, const float* H, int64_t ldh, void* stream) {
  GS_REQUIRE(rows >= 0 && groups >= 1 && cols >= 0 && D && H && ldh >= cols);
  const int64_t n = (int64_t)rows * groups * cols;
  if (n == 0) return GS_OK;
  relu_mask_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(rows, groups, cols, D, H, ldh);
  return finish_launch("relu_mask");
}

int gs_softmax_residual_f32(int32_t rows, int32_t C, const float* Z, int64_t ldz, const int32_t* label,
                            const float* row_scale, float* S, float* R, float* nll, void* stream) {
  GS_REQUIRE(rows >= 0 && C > 0 && Z && label && ldz >= C);
  if (rows == 0) return GS_OK;
  softmax_residual_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(rows, C, Z, ldz, label, row_scale, S, R, nll);
  return finish_launch("softmax_residual");
}

int gs_expand_class_blocks_f32(int32_t rows, int32_t C, int32_t nblk, const float* R, const int32_t* blk, float* E,
                               void* stream) {
  GS_REQUIRE(rows >= 0 && C > 0 && nblk > 0 && R && blk && E);
  const int64_t n = (int64_t)rows * nblk * C;
  if (n == 0) return GS_OK;
  expand_class_blocks_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(rows, C, nblk, R, blk, E);
  return finish_launch("expand_class_blocks");
}

int gs_pick_class_blocks_f32(int32_t rows, int32_t C, int32_t nblk, const float* Zf, const int32_t* blk, float* Q,
                             void* stream) {
  GS_REQUIRE(rows >= 0 && C > 0 && nblk > 0 && Zf && blk && Q);
  const int64_t n = (int64_t)rows * C;
  if (n == 0) return GS_OK;
  pick_class_blocks_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(rows, C, nblk, Zf, blk, Q);
  return finish_launch("pick_class_blocks");
}

int gs_softmax_jvp_f32(int32_t rows, int32_t C, const float* S, const float* Q, const float* row_scale, float* dZ,
                       void* stream) {
  GS_REQUIRE(rows >= 0 && C > 0 && S && Q && dZ);
  if (rows == 0) return GS_OK;
  softmax_jvp_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(rows, C, S, Q, row_scale, dZ);
  return finish_launch("softmax_jvp");
}

int gs_match_col_stats_f32(int32_t rows, int32_t cols, const float* gs_, const float* gr, int64_t ld, float* stats,
                           int64_t stats_ld, void* stream) {
  GS_REQUIRE(rows >= 0 && cols >= 0 && gs_ && gr && stats && ld >= cols);
  if (cols == 0) return GS_OK;
  match_col_stats_kernel<<<(cols + 31) / 32, 256, 0, as_stream(stream)>>>(rows, cols, gs_, gr, ld, stats, stats_ld);
  return finish_launch("match_col_stats");
}

int gs_match_finalize_f32(int metric, int32_t n_par, const int32_t* par_off, const int32_t* par_width,
                          const int32_t* par_is_bias, int32_t n_class, const float* coeff, const float* stats,
                          int64_t stats_ld, float* alpha, float* beta, float* class_loss, float* loss_out,
                          void* stream) {
  GS_REQUIRE(metric >= 0 && metric <= 2 && n_par > 0 && par_off && par_width && par_is_bias && n_class > 0 && coeff);
  GS_REQUIRE(stats && alpha && beta && class_loss && loss_out);
  float* scratch = class_loss;
  match_finalize_kernel<<<n_class, 256, 0, as_stream(stream)>>>(metric, n_par, par_off, par_width, par_is_bias, coeff,
                                                               stats, stats_ld, alpha, beta, scratch);
  int rc = finish_launch("match_finalize");
  if (rc) return rc;
  sum_small_kernel<<<1, 32, 0, as_stream(stream)>>>(n_class, scratch, loss_out);
  return finish_launch("sum_small");
}

int gs_match_apply_f32(int32_t rows, int32_t cols, const float* gs_, const float* gr, int64_t ld, const float* alpha,
                       const float* beta, float* G, void* stream) {
  GS_REQUIRE(rows >= 0 && cols >= 0 && gs_ && gr && alpha && beta && G && ld >= cols);
  const int64_t n = (int64_t)rows * cols;
  if (n == 0) return GS_OK;
  match_apply_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(rows, cols, gs_, gr, ld, alpha, beta, G);
  return finish_launch("match_apply");
}

int gs_segment_colsum_f32(int32_t G, const int32_t* seg, const int32_t* out_block, int32_t cols, const float* X,
                          int64_t ldx, int32_t n_rows, float* out, void* stream) {
  GS_REQUIRE(G >= 0 && seg && out_block && cols >= 0 && X && out && ldx >= cols && n_rows >= 0);
  if (G == 0 || cols == 0 || n_rows == 0) return GS_OK;
  dim3 grid((cols + 31) / 32, (n_rows + 8 * kColsumRun - 1) / (8 * kColsumRun));
  segment_colsum_kernel<<<grid, 256, 0, as_stream(stream)>>>(G, seg, out_block, cols, X, ldx, out);
  return finish_launch("segment_colsum");
}

int gs_dense_gcn_norm_fwd_f32(int32_t n, const float* A, float* Ahat, float* r, void* stream) {
  GS_REQUIRE(n > 0 && A && Ahat && r);
  dense_norm_rowsum_kernel<<<(n + 7) / 8, 256, 0, as_stream(stream)>>>(n, A, r);
  int rc = finish_launch("dense_norm_rowsum");
  if (rc) return rc;
  dense_norm_scale_kernel<<<blocks_for((int64_t)n * n), 256, 0, as_stream(stream)>>>(n, A, r, Ahat);
  return finish_launch("dense_norm_scale");
}

int gs_dense_gcn_norm_bwd_f32(int32_t n, const float* dAhat, const float* Ahat, const float* r, float* dA, float* work,
                              void* stream) {
  GS_REQUIRE(n > 0 && dAhat && Ahat && r && dA && work);
  dense_norm_bwd_dots_kernel<<<(n + 7) / 8, 256, 0, as_stream(stream)>>>(n, dAhat, Ahat, work);
  int rc = finish_launch("dense_norm_bwd_dots");
  if (rc) return rc;
  dense_norm_bwd_apply_kernel<<<blocks_for((int64_t)n * n), 256, 0, as_stream(stream)>>>(n, dAhat, r, work, dA);
  return finish_launch("dense_norm_bwd_apply");
}

int gs_adam_step_f32(int64_t n, float* p, const float* g, float* m, float* v, int32_t step, double lr, double beta1,
                     double beta2, double eps, void* stream) {
  GS_REQUIRE(n >= 0 && p && g && m && v && step >= 1);
  if (n == 0) return GS_OK;
  // scalar prologue in double exactly like torch's python floats (torch/optim/adam.py _single_tensor_adam)
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  adam_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(n, p, g, m, v, step_size, bc2_sqrt, (float)(1.0 - beta1),
                                                            (float)beta2, (float)(1.0 - beta2), (float)eps);
  return finish_launch("adam");
}

int gs_adam_table_f32(int32_t steps, double lr, double beta1, double beta2, float* table_host) {
  GS_REQUIRE(steps >= 0 && table_host);
  for (int32_t t = 0; t < steps; ++t) {   // the scalar prologue of gs_adam_step_f32 for step t + 1, same roundings
    const double bc1 = 1.0 - pow(beta1, (double)(t + 1));
    const double bc2 = 1.0 - pow(beta2, (double)(t + 1));
    table_host[2 * t] = (float)(lr / bc1);
    table_host[2 * t + 1] = (float)sqrt(bc2);
  }
  return GS_OK;
}

int gs_adam_step_table_f32(int64_t n, float* p, const float* g, float* m, float* v, const float* table,
                           const int32_t* step_dev, double beta1, double beta2, double eps, void* stream) {
  GS_REQUIRE(n >= 0 && p && g && m && v && table && step_dev);
  if (n == 0) return GS_OK;
  adam_table_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(n, p, g, m, v, table, step_dev, (float)(1.0 - beta1),
                                                                  (float)beta2, (float)(1.0 - beta2), (float)eps);
  return finish_launch("adam_table");
}

int gs_counter_add_i32(int32_t* counter, int32_t inc, void* stream) {
  GS_REQUIRE(counter);
  counter_add_kernel<<<1, 1, 0, as_stream(stream)>>>(counter, inc);
  return finish_launch("counter_add");
}

int gs_axpby_f32(int64_t n, float a, const float* x, float b, float* y, void* stream) {
  GS_REQUIRE(n >= 0 && x && y);
  if (n == 0) return GS_OK;
  axpby_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(n, a, x, b, y);
  return finish_launch("axpby");
}

}  // extern "C"
