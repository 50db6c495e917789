/* mkl.h stand-in for building the reference's gemm.c against OpenBLAS (oracle/build_ref.py build_blas): the CBLAS subset
 * Executable/gemm.c:82-89 uses.  TEST INFRASTRUCTURE - nothing in the product includes this. */
#ifndef SRT_ORACLE_MKL_SHIM_H
#define SRT_ORACLE_MKL_SHIM_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int M, int N, int K, float alpha,
                 const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc);
void openblas_set_num_threads(int n);
#endif
