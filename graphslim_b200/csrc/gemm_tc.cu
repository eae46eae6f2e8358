// tcgen05 / TMA dense contractions (filled in below); returns GS_ENOSYS for shapes it does not cover so
// gs_gemm_f32 falls through to the exact SIMT path.
#include "common.cuh"

namespace gs {
int gemm_tc_dispatch(int ta, int tb, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float beta, float* C, int64_t ldc, int precision, cudaStream_t st) {
  (void)ta; (void)tb; (void)M; (void)N; (void)K; (void)alpha; (void)A; (void)lda; (void)B; (void)ldb; (void)beta;
  (void)C; (void)ldc; (void)precision; (void)st;
  return GS_ENOSYS;
}
}  // namespace gs
