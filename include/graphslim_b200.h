/*
 * graphslim_b200 -- C ABI of the B200-native GCond hot path.
 *
 * The reference (Emory-Melody/GraphSlim, pure Python) has no FFI of its own; its hot path
 * bottoms out in third-party wheels (torch_sparse, ATen).  This header declares the
 * entry points a binding for that path would bind, each citing the reference call site it
 * replaces (paths relative to /root/reference/graphslim).  INTEGRATION.md shows the ctypes
 * stub that wires them under the reference's GCond/GCondX classes.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every device pointer is caller-owned HBM
 *   - matrices are row-major fp32 with an explicit leading dimension (elements)
 *   - CSR indices are int32, values fp32
 *   - `stream` is a cudaStream_t passed as void*; all launches are asynchronous on it
 *   - return value: 0 on success, a positive cudaError_t, or a negative GS_E* code
 *   - no hidden allocation, no global state
 */
#ifndef GRAPHSLIM_B200_H
#define GRAPHSLIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_OK 0
#define GS_EINVAL (-22)
#define GS_ENOSPC (-28)
#define GS_ENOSYS (-38)

/* library / device --------------------------------------------------------------------- */
int gs_version(void);
/* Number of kernels this library has launched since the last reset (bench.py's gpu_launches). */
int64_t gs_launch_count(void);
void gs_launch_count_reset(void);
const char* gs_last_error(void);

/* ---- sparse propagation ----------------------------------------------------------------
 * Y[r,:] (+)= sum_e val[e] * X[col[e],:]   for e in rowptr[r]..rowptr[r+1]
 * replaces torch_sparse.matmul / SparseTensor.__matmul__ at models/sgc.py:47,51 and
 * models/layers.py:41 (and the full-graph propagation of models/base.py:168-173).
 * With `col` holding global node ids it also fuses the `features[n_id]` gather of
 * condensation/gcond_base.py:214.  The backward w.r.t. the dense operand is the same call on
 * the transposed structure (A_hat is symmetric for the full graph; sampled blocks get their
 * transpose from gs_sample_step).
 * chunk_row/chunk_beg/chunk_end (may be NULL): slices (<= long_thr non-zeros each) of the rows whose
 * degree exceeds long_thr (power-law tails).  Those rows are cleared and their slices accumulated with
 * atomics; every other row is a plain warp-per-row item.  (graph_utils.build_row_chunks builds the list.) */
int gs_spmm_csr_f32(int32_t n_rows, const int32_t* rowptr, const int32_t* col, const float* val,
                    const float* X, int64_t ldx, int32_t F, float* Y, int64_t ldy, int accumulate,
                    int32_t n_chunks, int32_t long_thr, const int32_t* chunk_row, const int32_t* chunk_beg,
                    const int32_t* chunk_end, void* stream);

/* The same product run `tile_cols` feature columns at a time (a multiple of 4; the caller sizes n_src * tile_cols * 4 B
 * to fit the L2), for X larger than the L2: the gathered column slice stays L2-resident while the rows sweep it, at the
 * price of re-reading (col, val) once per slice.  Equal to gs_spmm_csr_f32 up to fp32 reassociation (bit-identical on full slices of >= 128 floats). */
int gs_spmm_csr_tiled_f32(int32_t n_rows, const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                          int64_t ldx, int32_t F, float* Y, int64_t ldy, int accumulate, int32_t n_chunks,
                          int32_t long_thr, const int32_t* chunk_row, const int32_t* chunk_beg,
                          const int32_t* chunk_end, int32_t tile_cols, void* stream);

/* Tuning of the wide (F >= 128) SpMM kernel; process-wide, not thread-safe against concurrent launches.
 *   impl   1 = one row per warp, (col,val) staged in shared memory (default: fastest at every measured point);
 *          2 = pipelined multi-row warps (kept as a measured negative result: registers cost too much occupancy)
 *   unr    feature-row gathers in flight per warp (0 = auto; impl 1: 4|8; impl 2: 2|4|8)
 *   group  work items per warp for impl 2 (0 = chosen from the problem size, else 1..32)
 *   flags  bit0 streaming stores of Y, bit1 evict-first loads of (col,val), bit2 (impl 1) 256-byte L2 prefetch on the
 *          sequential streams (rowptr, col, val)  -- none of them moved a measured point by more than 2 %
 *   wpb    impl 1: warps per CTA (0 = auto, 8|4|2)
 *   max_nv impl 1: widest column tile in units of 32 float4 (0 = auto = 8, 8|4|2|1); narrower tiles use fewer
 *          registers per warp and win on dense graphs with wide rows
 * Defaults: impl 1, everything else auto.
 * Every setting produces the same bits (CSR-order fmaf accumulation); no reference counterpart. */
int gs_spmm_set_tuning(int impl, int unr, int group, int flags, int wpb, int max_nv);

/* Transpose-free backward through a rectangular block: dX[col[e],:] += val[e] * dY[r,:] (atomics).
 * autograd of torch_sparse.matmul inside torch.autograd.grad at condensation/gcond_base.py:223. */
int gs_spmm_csr_scatter_f32(int32_t n_rows, const int32_t* rowptr, const int32_t* col, const float* val,
                            const float* dY, int64_t ldy, int32_t F, float* dX, int64_t ldx, void* stream);

/* out[i,:] = X[idx[i],:]   (features[n_id], condensation/gcond_base.py:214) */
int gs_gather_rows_f32(int32_t n, const int32_t* idx, const float* X, int64_t ldx, int32_t F, float* out,
                       int64_t ldo, void* stream);

/* Symmetric GCN normalisation of CSR values on the device, bit-exact with utils.py:451-458:
 * val_out[e] = (float)((r[row(e)] * (double)a[e]) * r[col[e]]) with r supplied in float64. */
int gs_csr_gcn_norm_f64(int32_t n_rows, const int32_t* rowptr, const int32_t* col, const float* a,
                        const double* r, float* val_out, void* stream);

/* ---- dense contractions ----------------------------------------------------------------
 * C = alpha * op(A) * op(B) + beta * C, row-major, op = transpose when t? != 0.
 * replaces the ATen matmuls of models/sgc.py:39,49, models/layers.py:40-46,377,
 * models/parametrized_adj.py:57-71 and their autograd.
 * precision: 0 = fp32 FMA (SIMT), 1 = tcgen05 3xBF16 split (fp32-class accuracy),
 *            2 = tcgen05 single BF16 (looser, stated in DESIGN.md).
 * workspace: device scratch of at least gs_gemm_workspace_bytes(M,N,K,precision) bytes (16-byte aligned) that
 *            receives the BF16 tile image of op(B); may be NULL when that returns 0.           */
int64_t gs_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K, int precision);
int gs_gemm_f32(int ta, int tb, int32_t M, int32_t N, int32_t K, float alpha, const float* A, int64_t lda,
                const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int precision, void* workspace,
                int64_t workspace_bytes, void* stream);

/* The same product with the element-wise tail of the layer fused into the store:
 *   C = epi(alpha * op(A) * op(B) + beta * C),  epi(v)[r,c] = mask ? (mask[r,c] > 0 ? w : 0) : w,
 *   w = relu ? max(v + bias[c], 0) : v + bias[c]        (bias / mask may be NULL)
 * i.e. `x @ W + b` followed by ReLU (models/layers.py:377-381, models/sgc.py:39-41) and, in the backward, the ReLU
 * mask of the saved activation.  The product is never K-split, so each output element is finished by one thread. */
int gs_gemm_epi_f32(int ta, int tb, int32_t M, int32_t N, int32_t K, float alpha, const float* A, int64_t lda,
                    const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int relu,
                    const float* mask, int64_t ldmask, int precision, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* Grouped K-segmented product for per-class weight gradients:
 * C[:, out_block[g]*N : (out_block[g]+1)*N] = A[seg[g]:seg[g+1], :M]^T * B[seg[g]:seg[g+1], :N]
 * (autograd.grad w.r.t. layer weights, one class per group; condensation/gcond_base.py:223,234).
 * K_total = rows of A and B.  precision 0: fp32 SIMT, C is overwritten.  precision 1/2: tcgen05; requires every
 * seg[g] to be a multiple of 64 (gs_sampler_set_align) and C zero-initialised by the caller (atomic accumulation);
 * workspace as for gs_gemm_f32 with (M, N, K_total).                                                         */
int gs_gemm_grouped_tn_f32(int32_t G, const int32_t* seg, const int32_t* out_block, int32_t M, int32_t N,
                           int32_t K_total, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                           int64_t ldc, int precision, void* workspace, int64_t workspace_bytes, void* stream);

/* out[out_block[g]*cols + c] += sum over rows seg[g]..seg[g+1] of X[r, c]  (per-class bias gradients,
 * autograd.grad w.r.t. the biases at condensation/gcond_base.py:223,234).  seg is non-decreasing, n_rows >= seg[G]
 * bounds the row range (rows of X); `out` is zero-initialised by the caller and accumulated with atomics. */
int gs_segment_colsum_f32(int32_t G, const int32_t* seg, const int32_t* out_block, int32_t cols, const float* X,
                          int64_t ldx, int32_t n_rows, float* out, void* stream);

/* ---- small fused kernels of the condense model ------------------------------------------- */
/* Z[r,c] += bias[c]; optional ReLU (models/layers.py:48-51,378-381; models/sgc.py:41) */
int gs_bias_act_f32(int32_t rows, int32_t cols, float* Z, int64_t ldz, const float* bias, int relu, void* stream);
/* D[r, g, c] *= (H[r,c] > 0) for g < groups  (ReLU backward, broadcast over the class axis) */
int gs_relu_mask_f32(int32_t rows, int32_t groups, int32_t cols, float* D, const float* H, int64_t ldh,
                     void* stream);
/* S = softmax(Z) rowwise; R = (S - onehot(label)) * row_scale; nll[r] = -log S[r,label]
 * (F.log_softmax + F.nll_loss and their gradient, models/sgc.py:57, gcond_base.py:221,228-231) */
int gs_softmax_residual_f32(int32_t rows, int32_t C, const float* Z, int64_t ldz, const int32_t* label,
                            const float* row_scale, float* S, float* R, float* nll, void* stream);
/* E[r, blk[r]*C + c] = R[r,c], zero elsewhere (E is rows x nblk*C) */
int gs_expand_class_blocks_f32(int32_t rows, int32_t C, int32_t nblk, const float* R, const int32_t* blk, float* E,
                               void* stream);
/* Q[r,c] = Zf[r, blk[r]*C + c] */
int gs_pick_class_blocks_f32(int32_t rows, int32_t C, int32_t nblk, const float* Zf, const int32_t* blk, float* Q,
                             void* stream);
/* dZ[r,:] = S[r,:] * (q - <S[r,:], q>),  q = Q[r,:] * row_scale[r]   (softmax Jacobian-vector product) */
int gs_softmax_jvp_f32(int32_t rows, int32_t C, const float* S, const float* Q, const float* row_scale, float* dZ,
                       void* stream);

/* ---- gradient matching (condensation/utils.py:12-106) ------------------------------------
 * Column statistics of two (rows x cols) gradient matrices: stats[0..3][col] =
 * <gs,gr>, |gs|^2, |gr|^2, |gs-gr|^2.                                                       */
int gs_match_col_stats_f32(int32_t rows, int32_t cols, const float* gs, const float* gr, int64_t ld, float* stats,
                           int64_t stats_ld, void* stream);
/* Turns the statistics of all parameters into the loss and the per-column coefficients
 * (alpha, beta) with  dLoss/dgs[:,j] = alpha[j]*gr[:,j] + beta[j]*gs[:,j].
 * metric: 0 'ours', 1 'mse', 2 'cos'.  par_off[p]..par_off[p+1] = columns of parameter p in the packed
 * stats; par_width[p] = per-class width; par_is_bias[p]; coeff[c] = n_c / N'.  loss_out += loss. */
int gs_match_finalize_f32(int metric, int32_t n_par, const int32_t* par_off, const int32_t* par_width,
                          const int32_t* par_is_bias, int32_t n_class, const float* coeff, const float* stats,
                          int64_t stats_ld, float* alpha, float* beta, float* class_loss /* n_class scratch */,
                          float* loss_out, void* stream);
/* G[:,j] = alpha[j]*gr[:,j] + beta[j]*gs[:,j] */
int gs_match_apply_f32(int32_t rows, int32_t cols, const float* gs, const float* gr, int64_t ld, const float* alpha,
                       const float* beta, float* G, void* stream);

/* ---- dense GCN normalisation (utils.py:429-439) and its backward --------------------------- */
int gs_dense_gcn_norm_fwd_f32(int32_t n, const float* A, float* Ahat, float* r, void* stream);
int gs_dense_gcn_norm_bwd_f32(int32_t n, const float* dAhat, const float* Ahat, const float* r, float* dA,
                              float* work /* 2n floats */, void* stream);

/* ---- PGE pairwise adjacency MLP (models/parametrized_adj.py:40-77) --------------------------
 * Pair k = i*n + j has input [x_j, x_i]; layer 1 is evaluated in split form Pa[j] + Pb[i].
 * Rows are processed in `nchunk` contiguous chunks with their own BatchNorm statistics
 * (np.array_split semantics, parametrized_adj.py:41-55); chunk_off has nchunk+1 entries.     */
/* BN1 batch statistics of Pa[j]+Pb[i] per chunk: mean/rstd (nchunk x h) */
int gs_pge_l1_stats_f32(int32_t n, int32_t h, const float* Pa, const float* Pb, int32_t nchunk,
                        const int64_t* chunk_off, float eps, float* mean, float* rstd,
                        double* work /* 2*nchunk*h */, void* stream);
/* Unchunked case (one BatchNorm batch over all n*n pairs): the statistics of Pa[j]+Pb[i] factorise exactly into the
 * column statistics of Pa and Pb (mean = mean_a + mean_b, var = var_a + var_b), 2n rows read instead of n*n.
 * col_mean (2 x h): column means of Pa and Pb, reused by gs_pge_bn1_bwd_closed_f32.  parametrized_adj.py:57-70. */
int gs_pge_l1_stats_closed_f32(int32_t n, int32_t h, const float* Pa, const float* Pb, float eps, float* mean,
                               float* rstd, float* col_mean, void* stream);
/* BN1 + ReLU backward of the factorised layer 1 in one pass over dH1 (unchunked): dPa[j] = sum_i dPre[(i,j)],
 * dPb[i] = sum_j dPre[(i,j)], dgamma1, dbeta1 (autograd of parametrized_adj.py:57-66).
 * work: 16-byte aligned, >= 16*h + 8*n*h bytes. */
int gs_pge_bn1_bwd_closed_f32(int32_t n, int32_t h, const float* dH1, const float* Pa, const float* Pb,
                              const float* mean, const float* rstd, const float* gamma, const float* beta,
                              const float* col_mean, float* dPa, float* dPb, float* dgamma, float* dbeta, void* work,
                              int64_t work_bytes, void* stream);
/* H1[k,:] = relu(gamma*(Pa[j]+Pb[i]-mean)*rstd + beta) */
int gs_pge_l1_expand_f32(int32_t n, int32_t h, const float* Pa, const float* Pb, int32_t nchunk,
                         const int64_t* chunk_off, const float* mean, const float* rstd, const float* gamma,
                         const float* beta, float* H1, void* stream);
/* per-chunk column mean / rstd of Y (rows x h), biased variance */
int gs_col_stats_chunked_f32(int64_t rows, int32_t h, const float* Y, int32_t nchunk, const int64_t* chunk_off,
                             float eps, float* mean, float* rstd, double* work /* 2*nchunk*h */, void* stream);
/* E[k] = relu(bn2(Y2[k,:])) . w3 + b3 */
int gs_pge_l3_f32(int64_t rows, int32_t h, const float* Y2, int32_t nchunk, const int64_t* chunk_off,
                  const float* mean, const float* rstd, const float* gamma, const float* beta, const float* w3,
                  const float* b3, float* E, void* stream);
/* A = sigmoid((E + E^T)/2) with zero diagonal */
int gs_pge_symm_sigmoid_f32(int32_t n, const float* E, float* A, void* stream);
/* dE = (T + T^T)/2, T = dA * A*(1-A) off-diagonal */
int gs_pge_symm_sigmoid_bwd_f32(int32_t n, const float* dA, const float* A, float* dE, void* stream);
/* layer-3 + BN2 backward statistics: per chunk s1 = sum dYhat, s2 = sum dYhat*xhat (overwritten);
 * dw3 (h) and db3 (1) are accumulated into.  work: 2*nchunk*h + h + 1 doubles. */
int gs_pge_l3_bwd_stats_f32(int64_t rows, int32_t h, const float* Y2, const float* dE, int32_t nchunk,
                            const int64_t* chunk_off, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, const float* w3, float* s1, float* s2, float* dw3, float* db3,
                            double* work, void* stream);
/* dY2[k,:] = gamma*rstd*(dYhat - s1/m - xhat*s2/m) */
int gs_pge_bn2_bwd_apply_f32(int64_t rows, int32_t h, const float* Y2, const float* dE, int32_t nchunk,
                             const int64_t* chunk_off, const float* mean, const float* rstd, const float* gamma,
                             const float* beta, const float* w3, const float* s1, const float* s2, float* dY2,
                             void* stream);
/* BN1 backward statistics from dH1 (rows x h): s1 = sum dYhat1, s2 = sum dYhat1*xhat1 per chunk */
int gs_pge_bn1_bwd_stats_f32(int32_t n, int32_t h, const float* dH1, const float* Pa, const float* Pb, int32_t nchunk,
                             const int64_t* chunk_off, const float* mean, const float* rstd, const float* gamma,
                             const float* beta, float* s1, float* s2, double* work /* 2*nchunk*h */, void* stream);
/* dPa[j,:] = sum_i dY1[i,j,:],  dPb[i,:] = sum_j dY1[i,j,:] */
int gs_pge_bn1_bwd_reduce_f32(int32_t n, int32_t h, const float* dH1, const float* Pa, const float* Pb, int32_t nchunk,
                              const int64_t* chunk_off, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const float* s1, const float* s2, float* dPa, float* dPb,
                              void* stream);

/* ---- row-sharded PGE (one box, pair rows (i, j) with i in this rank's slice; graphslim_b200/pge.py ShardedPGE).
 * The kernels are the ones above; every reduction over all N'^2 pair rows is cut into a per-rank partial and a
 * replicated combine with one small collective in between (no reference counterpart: the reference is single-GPU).
 *   gs_pge_l1_expand_rows_f32     H1 rows of the slice: n_i x n pair rows from Pa (all j) and Pb_rows (the slice's i)
 *   gs_col_stats_partial_f64      work[0..h) = sum(y - y_row0), work[h..2h) = sum((y - y_row0)^2) over the slice
 *   gs_col_stats_combine_f32      parts[world][3h] = (S1 | S2 | shift row) + counts[world] -> mean, rstd over all rows
 *   gs_pge_bn1_bwd_pass_rows_f32  linear reductions of dH1 over the slice into work = [t1,t2 (2h doubles) | Ga (n x h) |
 *                                 Gb (n x h, only the slice's rows non-zero)]; summed over ranks by the caller
 *   gs_pge_bn1_bwd_final_f32      dPa, dPb, dgamma1, dbeta1 from the summed work */
int gs_pge_l1_expand_rows_f32(int32_t n_i, int32_t n, int32_t h, const float* Pa, const float* Pb_rows,
                              const int64_t* chunk_off_rows, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, float* H1, void* stream);
int gs_col_stats_partial_f64(int64_t rows, int32_t h, const float* Y, const int64_t* chunk_off_rows, double* work,
                             void* stream);
int gs_col_stats_combine_f32(int32_t world, int32_t h, const double* parts, const int64_t* counts, float eps,
                             float* mean, float* rstd, void* stream);
int64_t gs_pge_bn1_bwd_work_bytes(int32_t n, int32_t h);
int gs_pge_bn1_bwd_pass_rows_f32(int32_t n_i, int32_t i_first, int32_t n, int32_t h, const float* dH1_rows,
                                 const float* Pa, const float* Pb, const float* mean, const float* rstd,
                                 const float* gamma, const float* beta, void* work, int64_t work_bytes, void* stream);
int gs_pge_bn1_bwd_final_f32(int32_t n, int32_t h, const float* Pa, const float* Pb, const float* rstd,
                             const float* gamma, const float* col_mean, const void* work, float* dPa, float* dPb,
                             float* dgamma, float* dbeta, void* stream);

/* ---- grouped TN products with MN-major, TMA-fed operands (csrc/grouped_tn.cu): the per-class weight gradients of the
 * real side, autograd.grad(loss_real, params) of condensation/gcond_base.py:221-224 for all classes at once.
 *   gs_gemm_grouped_mn_supported  1 when the shape is covered (M in {128, 256}; N in {128, 256} or N <= 64 with N % 4 == 0;
 *                                 precision 1 or 2), else use gs_gemm_grouped_tn_f32
 *   gs_gemm_grouped_mn_f32        C[:, out_block[g]*N : +N] += A[seg[g]:seg[g+1]]^T B[seg[g]:seg[g+1]] (C zeroed by the
 *                                 caller, segments 64-aligned); A, B row-major fp32 read by cp.async.bulk.tensor and
 *                                 converted to BF16 hi/lo in place in shared memory, no packing or transposition in HBM
 *   gs_mlp_bwd_grouped_f32        backward of the hidden ReLU layer of the 2-layer condense models (models/sgc.py:37-57,
 *                                 layers.py:36-51 under autograd): dA1 = (dU W2^T) . [H1 > 0] generated per k-stage on
 *                                 chip, gW1 += X^T dA1 and gb1 += colsum(dA1) per class; dA1 never reaches HBM */
int gs_gemm_grouped_mn_supported(int32_t M, int32_t N, int precision);
int gs_gemm_grouped_mn_f32(int32_t G, const int32_t* seg, const int32_t* out_block, int32_t M, int32_t N,
                           int32_t total_rows, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                           int64_t ldc, int precision, void* stream);
int gs_mlp_bwd_grouped_f32(int32_t G, const int32_t* seg, const int32_t* out_block, int32_t M, int32_t N, int32_t Cw,
                           int32_t total_rows, const float* X, int64_t ldx, const float* H1, int64_t ldh,
                           const float* dU, int64_t ldu, const float* W2, int64_t ldw2, float* gW1, int64_t ldc,
                           float* gb1, int precision, void* stream);

/* ---- fused PGE layer-2 pipeline (csrc/pge_fused.cu): tcgen05 + TMEM products whose N'^2 x h operands are produced on
 * chip, replacing gs_pge_l1_expand + gs_gemm + gs_col_stats (forward) and gs_pge_bn2_bwd_apply + two gs_gemm +
 * gs_pge_bn1_bwd_pass (backward) of models/parametrized_adj.py:57-71 and its autograd.  Unchunked BatchNorm, h = 128 or
 * 256, precision 1 (3xBF16) or 2 (BF16); pair rows (i, j) with i in [i_first, i_first + n_i) (a rank's slice), all j.
 *   gs_pge_fused_l2_fwd_f32     Y2 (n_i*n x h) = relu(bn1(Pa[j] + Pb[i])) W2^T, H1 generated in the A-producer;
 *                               stats = [sum y | sum y^2] (2h doubles, overwritten) from the epilogue
 *   gs_pge_stats_finalize_f32   mean / rstd of BatchNorm 2 from the (all-reduced) sums and the global row count
 *   gs_pge_fused_l2_bwd_dx_f32  dH1 = dY2 W2 with dY2 = bn2'(relu'(bn2(Y2)) dE w3) computed from the TMA-loaded raw Y2
 *                               tile; dH1 != NULL stores it, else it is masked by H1 > 0 and reduced into Ga[j] = sum_i,
 *                               Gb[i] = sum_j (n x h each, accumulated into: the caller zeroes them)
 *   gs_pge_fused_l2_bwd_dw_f32  dW2 (h x h, overwritten) = dY2^T H1, both operands produced on chip (MN-major UMMA)
 *   gs_pge_bn1_tsum_f64         adds t1 = sum g, t2 = sum g*xhat implied by Ga / Gb into work = [t1 | t2 | Ga | Gb]
 *                               (layout of gs_pge_bn1_bwd_pass_rows_f32), ready for gs_pge_bn1_bwd_final_f32
 * workspace: gs_pge_fused_workspace_bytes(h, precision) bytes (BF16 tile image of W2), 16-byte aligned. */
int64_t gs_pge_fused_workspace_bytes(int32_t h, int precision);
int gs_pge_fused_l2_fwd_f32(int32_t n, int32_t n_i, int32_t i_first, int32_t h, const float* Pa, const float* Pb,
                            const float* mean1, const float* rstd1, const float* gamma1, const float* beta1,
                            const float* W2, int64_t ldw, float* Y2, double* stats, int precision, void* workspace,
                            int64_t workspace_bytes, void* stream);
int gs_pge_stats_finalize_f32(int32_t h, const double* stats, double count, float eps, float* mean, float* rstd,
                              void* stream);
int gs_pge_fused_l2_bwd_dx_f32(int32_t n, int32_t n_i, int32_t i_first, int32_t h, const float* Pa, const float* Pb,
                               const float* mean1, const float* rstd1, const float* gamma1, const float* beta1,
                               const float* W2, int64_t ldw, const float* Y2, const float* dE, const float* mean2,
                               const float* rstd2, const float* gamma2, const float* beta2, const float* w3,
                               const float* s1, const float* s2, double count, float* Ga, float* Gb, float* dH1,
                               int precision, void* workspace, int64_t workspace_bytes, void* stream);
int gs_pge_fused_l2_bwd_dw_f32(int32_t n, int32_t n_i, int32_t i_first, int32_t h, const float* Pa, const float* Pb,
                               const float* mean1, const float* rstd1, const float* gamma1, const float* beta1,
                               const float* Y2, const float* dE, const float* mean2, const float* rstd2,
                               const float* gamma2, const float* beta2, const float* w3, const float* s1,
                               const float* s2, double count, float* dW2, int precision, void* stream);
int gs_pge_bn1_tsum_f64(int32_t n, int32_t h, const float* Pa, const float* Pb, const float* col_mean,
                        const float* rstd1, void* work, void* stream);

/* ---- optimiser (torch.optim.Adam defaults; condensation/gcond_base.py:68-69, gcond.py:44) ---- */
int gs_adam_step_f32(int64_t n, float* p, const float* g, float* m, float* v, int32_t step, double lr, double beta1,
                     double beta2, double eps, void* stream);
/* The same step with its two step-dependent scalars read from device memory, so that a launch no longer depends on
 * the step number and an inner-loop iteration (gcond.py:63-72) can be captured in a CUDA graph:
 *   gs_adam_table_f32 fills table_host[2*t], table_host[2*t+1] (t = 0..steps-1) with the step_size and sqrt(bias
 *   correction 2) that gs_adam_step_f32 computes for step t+1 (same double arithmetic, same roundings);
 *   gs_adam_step_table_f32 applies step *step_dev (zero-based, device) from the uploaded table;
 *   gs_counter_add_i32 advances the device step counter. */
int gs_adam_table_f32(int32_t steps, double lr, double beta1, double beta2, float* table_host);
int gs_adam_step_table_f32(int64_t n, float* p, const float* g, float* m, float* v, const float* table,
                           const int32_t* step_dev, double beta1, double beta2, double eps, void* stream);
int gs_counter_add_i32(int32_t* counter, int32_t inc, void* stream);
/* A recorded list of small dense operations executed by ONE persistent cooperative kernel (csrc/chain.cu): the
 * condense-model training step of the inner loop (graphslim/condensation/gcond.py:63-72 -- model.forward, F.nll_loss,
 * loss.backward, optimizer_model.step: ~35 dependent launches of 3-20 us each on at most N' rows) becomes one launch.
 * `ops_dev` is a device array; the kernel walks it in order with a grid barrier before every operation whose
 * `sync_before` is set (the recorder clears it when an operation touches nothing the operations since the last barrier
 * wrote or read-then-overwrite).  Exact fp32 FMA, fixed summation orders (bit-reproducible).  kinds:
 *   0 GEMM              C = epi(alpha op(A) op(B) + beta C)  (M,N,K, ta,tb, lda,ldb,ldc; bias[N], relu, mask[M x N] ldmask)
 *   1 SOFTMAX_RESIDUAL  A = Z (M rows x N classes, lda), B = int32 labels, bias = row scale or 0, C = S or 0, p5 = R
 *   2 COLSUM            C[N] = column sums of A (M x N, lda)
 *   3 ADAM_TABLE        gs_adam_step_table_f32: C = p, A = g, p5 = m, p6 = v, B = table, p7 = step counter, lda = n,
 *                       alpha = 1-beta1, beta = beta2, f0 = 1-beta2, f1 = eps
 *   4 COUNTER_ADD       *(int32*)C += M
 *   5 FILL              C[0..lda) = alpha */
typedef struct gs_chain_op {
  int32_t kind, ta, tb, relu;
  int32_t M, N, K, sync_before;
  int64_t lda, ldb, ldc, ldmask;
  float alpha, beta, f0, f1;
  uint64_t A, B, C, bias, mask, p5, p6, p7;   /* device addresses */
} gs_chain_op;
int gs_chain_run_f32(const gs_chain_op* ops_dev, int32_t n_ops, void* stream);
/* y = a*x + b*y */
int gs_axpby_f32(int64_t n, float a, const float* x, float b, float* y, void* stream);

/* ---- host side: bit-exact class / neighbour selection (dataset/loader.py:187-224) ------------
 * One call samples the blocks of every class of one outer step, drawing from an mt19937 state in the
 * layout of torch's CPU generator so the random stream interleaves exactly like the reference's.  */
/* `np.random.permutation(members of class c)[:batch]` for every class in order (graphslim/dataset/loader.py:222) on
 * numpy's legacy MT19937 stream: `key` (624 words) and `pos` are np.random.get_state()[1:3] and are advanced in place
 * (hand them back with np.random.set_state).  members: concatenated int64 ids, member_off: n_class + 1 offsets;
 * out: int32 batches back to back, out_off: n_class + 1 offsets.  Host only; runs without the interpreter lock. */
int gs_np_legacy_class_batches(uint32_t* key, int32_t* pos_io, int32_t n_class, const int64_t* members,
                               const int64_t* member_off, int32_t batch, int32_t* out, int32_t* out_off);
typedef struct gs_sampler gs_sampler;
gs_sampler* gs_sampler_create(int32_t n_nodes, const int64_t* rowptr, const int32_t* col, const float* val,
                              int32_t n_hops, const int32_t* fanout);
void gs_sampler_destroy(gs_sampler* s);
/* optional: int32 label per node; the labels of the target rows are then emitted with every step */
void gs_sampler_set_labels(gs_sampler* s, const int32_t* labels);
/* pad every class segment of every level to a multiple of `align` rows (pad rows: no edges, zero loss weight);
 * 64 makes the segments tile-aligned for the tensor-core grouped products */
void gs_sampler_set_align(gs_sampler* s, int32_t align);
/* worker threads for the last (largest) hop; default = min(hardware threads, 16); results do not depend on it */
void gs_sampler_set_threads(gs_sampler* s, int32_t n);
/* batch: concatenated class batches (node ids), batch_off[n_class+1]; materialise[c] != 0 selects the classes
 * whose blocks are written (others only advance the generator).  out: packed buffer of `out_cap` bytes;
 * desc: int64[64] table describing where each array landed.  Returns bytes used or a negative code. */
int64_t gs_sampler_sample_step(gs_sampler* s, int32_t n_class, const int64_t* batch, const int64_t* batch_off,
                               const uint8_t* materialise, uint32_t* mt_state, int32_t* mt_left, int32_t* mt_next,
                               uint8_t* out, int64_t out_cap, int64_t* desc);

/* The same step in two calls so a caller can pipeline them on two threads: begin_step owns the random stream (all
 * hops but the last + per-class segmentation of the last hop's stream, serial); finish_step samples the last hop of
 * every class in parallel, packs, and frees the job.  finish_step(t) may overlap begin_step(t+1).                */
typedef struct gs_sample_job gs_sample_job;
gs_sample_job* gs_sampler_begin_step(gs_sampler* s, int32_t n_class, const int64_t* batch, const int64_t* batch_off,
                                     const uint8_t* materialise, uint32_t* mt_state, int32_t* mt_left,
                                     int32_t* mt_next);
int64_t gs_sampler_finish_step(gs_sampler* s, gs_sample_job* job, uint8_t* out, int64_t out_cap, int64_t* desc);

/* ---- device-side class-batch neighbour sampler ---------------------------------------------
 * Same contract as gs_sampler_* above (TransAndInd.retrieve_class_sampler, dataset/loader.py:187-224, with
 * torch_geometric NeighborSampler / torch_sparse sample_adj under it), executed on the GPU against the CSR that is
 * already in HBM: bit-identical class blocks for the same torch CPU generator state, no host sampling and no H2D copy
 * of the blocks.  csrc/device_sampler.cu explains how the serial mt19937 stream and the unordered_set iteration order
 * are reproduced in parallel.  All pointers are DEVICE pointers unless noted; work is queued on `stream`. */
typedef struct gs_dsampler gs_dsampler;
gs_dsampler* gs_dsampler_create(int32_t n_nodes, const int32_t* d_rowptr, const int32_t* d_col, const float* d_val,
                                const int32_t* d_labels /* may be NULL */, int32_t n_hops,
                                const int32_t* fanout /* host */, int32_t n_class_max, int32_t batch_max,
                                int32_t align, void* stream);
void gs_dsampler_destroy(gs_dsampler* s);
/* bytes the packed output of one step can need / bytes of device scratch the handle owns */
int64_t gs_dsampler_out_capacity(const gs_dsampler* s);
int64_t gs_dsampler_scratch_bytes(const gs_dsampler* s);
/* torch's mt19937 engine (624 state words, `left`, `next`; host pointers) <-> the device-resident generator.
 * Both calls synchronise `stream`. */
int gs_dsampler_set_rng(gs_dsampler* s, const uint32_t* state, int32_t left, int32_t next, void* stream);
int gs_dsampler_get_rng(gs_dsampler* s, uint32_t* state, int32_t* left, int32_t* next, void* stream);
/* One outer step for n_class classes: d_batch = concatenated class batches (int32 node ids), d_batch_off = n_class+1
 * offsets, d_materialise = per-class 0/1 or NULL (class sharding: skipped classes still consume their random draws),
 * max_batch = largest class batch (host value).  d_desc (64 x int64) receives the layout of d_out exactly as
 * gs_sampler_finish_step's `desc`, plus [60] draws consumed, [62] bytes used, [63] 0 or GS_ENOSPC. */
int gs_dsampler_sample_step(gs_dsampler* s, int32_t n_class, const int32_t* d_batch, const int32_t* d_batch_off,
                            const uint8_t* d_materialise, int32_t max_batch, uint8_t* d_out, int64_t out_cap,
                            int64_t* d_desc, void* stream);
/* phase cycle counters of the serial sampling kernel (all zero unless the library was built with -DGS_DS_PROFILE) */
int gs_dsampler_debug_counters(gs_dsampler* s, int64_t* out16 /* host */, void* stream);
/* host-side check of the std::unordered_set iteration-order restatement used by the device sampler (tests) */
int64_t gs_uset_emul_order(const int64_t* keys, int64_t n, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif
