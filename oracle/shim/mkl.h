/*
 * oracle/shim/mkl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Declaration-only stand-in for Intel MKL's <mkl.h> so that the UNMODIFIED
 * reference sources under /root/reference compile in a container that has no
 * MKL (see oracle/Makefile).  Only the names the reference actually uses are
 * declared (reference: include/types.h:33-78 FP* aliases, include/matUtils.h,
 * src/sparseMatrix.cpp, src/denseMatrix.cpp).  Definitions live in
 * mkl_shim.cpp: BLAS/LAPACK forward to the ILP64 OpenBLAS shipped inside
 * numpy.libs, the MKL-only sparse routines are plain OpenMP loops.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef MKL_ILP64
typedef long long MKL_INT;
typedef unsigned long long MKL_UINT;
#else
typedef int MKL_INT;
typedef unsigned int MKL_UINT;
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef CBLAS_LAYOUT CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;

/* ---- CBLAS (ILP64) ---- */
void cblas_sgemm(const CBLAS_LAYOUT, const CBLAS_TRANSPOSE, const CBLAS_TRANSPOSE,
                 const MKL_INT m, const MKL_INT n, const MKL_INT k, const float alpha,
                 const float *a, const MKL_INT lda, const float *b, const MKL_INT ldb,
                 const float beta, float *c, const MKL_INT ldc);
void cblas_sgemv(const CBLAS_LAYOUT, const CBLAS_TRANSPOSE, const MKL_INT m, const MKL_INT n,
                 const float alpha, const float *a, const MKL_INT lda, const float *x,
                 const MKL_INT incx, const float beta, float *y, const MKL_INT incy);
void cblas_ssymv(const CBLAS_LAYOUT, const CBLAS_UPLO, const MKL_INT n, const float alpha,
                 const float *a, const MKL_INT lda, const float *x, const MKL_INT incx,
                 const float beta, float *y, const MKL_INT incy);
float cblas_sdot(const MKL_INT n, const float *x, const MKL_INT incx, const float *y,
                 const MKL_INT incy);
float cblas_sasum(const MKL_INT n, const float *x, const MKL_INT incx);
float cblas_snrm2(const MKL_INT n, const float *x, const MKL_INT incx);
void cblas_saxpy(const MKL_INT n, const float a, const float *x, const MKL_INT incx, float *y,
                 const MKL_INT incy);
void cblas_sscal(const MKL_INT n, const float a, float *x, const MKL_INT incx);
void cblas_scopy(const MKL_INT n, const float *x, const MKL_INT incx, float *y,
                 const MKL_INT incy);
size_t cblas_isamin(const MKL_INT n, const float *x, const MKL_INT incx);

/* ---- MKL sparse BLAS / service (deprecated NIST-style interface) ---- */
void mkl_scsrmm(const char *transa, const MKL_INT *m, const MKL_INT *n, const MKL_INT *k,
                const float *alpha, const char *matdescra, const float *val,
                const MKL_INT *indx, const MKL_INT *pntrb, const MKL_INT *pntre,
                const float *b, const MKL_INT *ldb, const float *beta, float *c,
                const MKL_INT *ldc);
void mkl_scscmm(const char *transa, const MKL_INT *m, const MKL_INT *n, const MKL_INT *k,
                const float *alpha, const char *matdescra, const float *val,
                const MKL_INT *indx, const MKL_INT *pntrb, const MKL_INT *pntre,
                const float *b, const MKL_INT *ldb, const float *beta, float *c,
                const MKL_INT *ldc);
void mkl_scsrcsc(const MKL_INT *job, const MKL_INT *n, float *acsr, MKL_INT *ja, MKL_INT *ia,
                 float *acsc, MKL_INT *ja1, MKL_INT *ia1, MKL_INT *info);
void mkl_sdnscsr(const MKL_INT *job, const MKL_INT *m, const MKL_INT *n, float *adns,
                 const MKL_INT *lda, float *acsr, MKL_INT *ja, MKL_INT *ia, MKL_INT *info);
void mkl_cspblas_scsrgemv(const char *transa, const MKL_INT *m, const float *a,
                          const MKL_INT *ia, const MKL_INT *ja, const float *x, float *y);
void mkl_somatcopy(const char ordering, const char trans, size_t rows, size_t cols,
                   const float alpha, const float *A, size_t lda, float *B, size_t ldb);
int mkl_set_num_threads_local(int nt);
void *mkl_malloc(size_t size, int align);
void mkl_free(void *p);

/* LAPACKE name referenced only through an unused macro (include/types.h:38). */
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
MKL_INT LAPACKE_sgesvd(int layout, char jobu, char jobvt, MKL_INT m, MKL_INT n, float *a,
                       MKL_INT lda, float *s, float *u, MKL_INT ldu, float *vt, MKL_INT ldvt,
                       float *superb);

#ifdef __cplusplus
}
#endif
