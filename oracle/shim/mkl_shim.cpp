/*
 * oracle/shim/mkl_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Definitions for the 24 numeric symbols the unmodified reference objects
 * leave undefined when built without Intel MKL:
 *   - CBLAS (ILP64) and the Fortran BLAS/LAPACK names Armadillo emits
 *     (ARMA_BLAS_LONG_LONG => 64-bit ints) forward to the ILP64 OpenBLAS
 *     0.3.30 inside numpy.libs (symbols scipy_cblas_*64_ / scipy_*_64_);
 *   - the MKL-only sparse/service routines are straightforward loops that
 *     implement exactly the one mode the reference uses:
 *       mkl_scsrmm  : transa='N', matdescra = {'G',.,.,'C',..}  (general,
 *                     zero-based, row-major dense operands)
 *                     call sites: include/matUtils.h:329, src/sparseMatrix.cpp:1776,1257
 *       mkl_scsrcsc : job = {1,0,0,0,0,1} CSC->CSR (include/matUtils.h:99-106;
 *                     unreachable at run time, see SURVEY Q3, kept so it links)
 *       mkl_somatcopy('C','T',..)  src/sparseMatrix.cpp:1513, src/infer.cpp:319
 *       mkl_sdnscsr : dense->CSR, job[0]=0 (src/denseMatrix.cpp:238)
 *       mkl_cspblas_scsrgemv : zero-based CSR y=A*x (include/matUtils.h:374-407, Spectra path)
 *   The sparse loops are OpenMP-parallel over rows so the oracle is a fair
 *   multi-core CPU baseline ("reference C++ over OpenBLAS + shim", not MKL).
 */
#include "mkl.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <omp.h>

typedef long long blasint;

extern "C" {
/* ---- OpenBLAS ILP64 entry points (numpy.libs/libscipy_openblas64_) ---- */
void scipy_cblas_sgemm64_(int, int, int, blasint, blasint, blasint, float, const float *, blasint,
                          const float *, blasint, float, float *, blasint);
void scipy_cblas_sgemv64_(int, int, blasint, blasint, float, const float *, blasint,
                          const float *, blasint, float, float *, blasint);
void scipy_cblas_ssymv64_(int, int, blasint, float, const float *, blasint, const float *, blasint,
                          float, float *, blasint);
float scipy_cblas_sdot64_(blasint, const float *, blasint, const float *, blasint);
float scipy_cblas_sasum64_(blasint, const float *, blasint);
float scipy_cblas_snrm264_(blasint, const float *, blasint);
void scipy_cblas_saxpy64_(blasint, float, const float *, blasint, float *, blasint);
void scipy_cblas_sscal64_(blasint, float, float *, blasint);
void scipy_cblas_scopy64_(blasint, const float *, blasint, float *, blasint);
size_t scipy_cblas_isamin64_(blasint, const float *, blasint);
void scipy_openblas_set_num_threads64_(int);
int scipy_openblas_get_num_threads64_(void);

void scipy_sgemm_64_(const char *, const char *, const blasint *, const blasint *, const blasint *,
                     const float *, const float *, const blasint *, const float *, const blasint *,
                     const float *, float *, const blasint *);
void scipy_dgemm_64_(const char *, const char *, const blasint *, const blasint *, const blasint *,
                     const double *, const double *, const blasint *, const double *,
                     const blasint *, const double *, double *, const blasint *);
void scipy_sgemv_64_(const char *, const blasint *, const blasint *, const float *, const float *,
                     const blasint *, const float *, const blasint *, const float *, float *,
                     const blasint *);
void scipy_dgemv_64_(const char *, const blasint *, const blasint *, const double *,
                     const double *, const blasint *, const double *, const blasint *,
                     const double *, double *, const blasint *);
float scipy_sdot_64_(const blasint *, const float *, const blasint *, const float *,
                     const blasint *);
double scipy_ddot_64_(const blasint *, const double *, const blasint *, const double *,
                      const blasint *);
void scipy_ssyrk_64_(const char *, const char *, const blasint *, const blasint *, const float *,
                     const float *, const blasint *, const float *, float *, const blasint *);
void scipy_ssyev_64_(const char *, const char *, const blasint *, float *, const blasint *, float *,
                     float *, const blasint *, blasint *);
void scipy_ssyevd_64_(const char *, const char *, const blasint *, float *, const blasint *,
                      float *, float *, const blasint *, blasint *, const blasint *, blasint *);

/* ---- CBLAS forwards ---- */
void cblas_sgemm(const CBLAS_LAYOUT l, const CBLAS_TRANSPOSE ta, const CBLAS_TRANSPOSE tb,
                 const MKL_INT m, const MKL_INT n, const MKL_INT k, const float alpha,
                 const float *a, const MKL_INT lda, const float *b, const MKL_INT ldb,
                 const float beta, float *c, const MKL_INT ldc)
{
    scipy_cblas_sgemm64_(l, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void cblas_sgemv(const CBLAS_LAYOUT l, const CBLAS_TRANSPOSE t, const MKL_INT m, const MKL_INT n,
                 const float alpha, const float *a, const MKL_INT lda, const float *x,
                 const MKL_INT incx, const float beta, float *y, const MKL_INT incy)
{
    scipy_cblas_sgemv64_(l, t, m, n, alpha, a, lda, x, incx, beta, y, incy);
}
void cblas_ssymv(const CBLAS_LAYOUT l, const CBLAS_UPLO u, const MKL_INT n, const float alpha,
                 const float *a, const MKL_INT lda, const float *x, const MKL_INT incx,
                 const float beta, float *y, const MKL_INT incy)
{
    scipy_cblas_ssymv64_(l, u, n, alpha, a, lda, x, incx, beta, y, incy);
}
float cblas_sdot(const MKL_INT n, const float *x, const MKL_INT incx, const float *y,
                 const MKL_INT incy)
{
    return scipy_cblas_sdot64_(n, x, incx, y, incy);
}
float cblas_sasum(const MKL_INT n, const float *x, const MKL_INT incx)
{
    return scipy_cblas_sasum64_(n, x, incx);
}
float cblas_snrm2(const MKL_INT n, const float *x, const MKL_INT incx)
{
    return scipy_cblas_snrm264_(n, x, incx);
}
void cblas_saxpy(const MKL_INT n, const float a, const float *x, const MKL_INT incx, float *y,
                 const MKL_INT incy)
{
    scipy_cblas_saxpy64_(n, a, x, incx, y, incy);
}
void cblas_sscal(const MKL_INT n, const float a, float *x, const MKL_INT incx)
{
    scipy_cblas_sscal64_(n, a, x, incx);
}
void cblas_scopy(const MKL_INT n, const float *x, const MKL_INT incx, float *y,
                 const MKL_INT incy)
{
    scipy_cblas_scopy64_(n, x, incx, y, incy);
}
size_t cblas_isamin(const MKL_INT n, const float *x, const MKL_INT incx)
{
    return scipy_cblas_isamin64_(n, x, incx);
}

/* ---- Fortran BLAS/LAPACK forwards (Armadillo, 64-bit ints) ---- */
void sgemm_(const char *ta, const char *tb, const blasint *m, const blasint *n, const blasint *k,
            const float *alpha, const float *a, const blasint *lda, const float *b,
            const blasint *ldb, const float *beta, float *c, const blasint *ldc)
{
    scipy_sgemm_64_(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void dgemm_(const char *ta, const char *tb, const blasint *m, const blasint *n, const blasint *k,
            const double *alpha, const double *a, const blasint *lda, const double *b,
            const blasint *ldb, const double *beta, double *c, const blasint *ldc)
{
    scipy_dgemm_64_(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void sgemv_(const char *t, const blasint *m, const blasint *n, const float *alpha, const float *a,
            const blasint *lda, const float *x, const blasint *incx, const float *beta, float *y,
            const blasint *incy)
{
    scipy_sgemv_64_(t, m, n, alpha, a, lda, x, incx, beta, y, incy);
}
void dgemv_(const char *t, const blasint *m, const blasint *n, const double *alpha,
            const double *a, const blasint *lda, const double *x, const blasint *incx,
            const double *beta, double *y, const blasint *incy)
{
    scipy_dgemv_64_(t, m, n, alpha, a, lda, x, incx, beta, y, incy);
}
float sdot_(const blasint *n, const float *x, const blasint *incx, const float *y,
            const blasint *incy)
{
    return scipy_sdot_64_(n, x, incx, y, incy);
}
double ddot_(const blasint *n, const double *x, const blasint *incx, const double *y,
             const blasint *incy)
{
    return scipy_ddot_64_(n, x, incx, y, incy);
}
void ssyrk_(const char *uplo, const char *trans, const blasint *n, const blasint *k,
            const float *alpha, const float *a, const blasint *lda, const float *beta, float *c,
            const blasint *ldc)
{
    scipy_ssyrk_64_(uplo, trans, n, k, alpha, a, lda, beta, c, ldc);
}
void ssyev_(const char *jobz, const char *uplo, const blasint *n, float *a, const blasint *lda,
            float *w, float *work, const blasint *lwork, blasint *info)
{
    scipy_ssyev_64_(jobz, uplo, n, a, lda, w, work, lwork, info);
}
void ssyevd_(const char *jobz, const char *uplo, const blasint *n, float *a, const blasint *lda,
             float *w, float *work, const blasint *lwork, blasint *iwork, const blasint *liwork,
             blasint *info)
{
    scipy_ssyevd_64_(jobz, uplo, n, a, lda, w, work, lwork, iwork, liwork, info);
}

/* ---- MKL-only routines: reference loops ---- */

/* C(m x n, row-major, ldc) = alpha * A(m x k CSR, zero-based) * B(k x n, row-major, ldb) + beta*C */
void mkl_scsrmm(const char *transa, const MKL_INT *m_, const MKL_INT *n_, const MKL_INT *k_,
                const float *alpha_, const char *matdescra, const float *val,
                const MKL_INT *indx, const MKL_INT *pntrb, const MKL_INT *pntre, const float *b,
                const MKL_INT *ldb_, const float *beta_, float *c, const MKL_INT *ldc_)
{
    if ((*transa != 'N' && *transa != 'n') || matdescra[0] != 'G' || matdescra[3] != 'C') {
        std::abort(); /* mode never used by the reference */
    }
    const MKL_INT m = *m_, n = *n_, ldb = *ldb_, ldc = *ldc_;
    const float alpha = *alpha_, beta = *beta_;
    (void)k_;
#pragma omp parallel for schedule(dynamic, 256)
    for (MKL_INT r = 0; r < m; ++r) {
        float *crow = c + (size_t)r * (size_t)ldc;
        if (beta == 0.0f) {
            for (MKL_INT j = 0; j < n; ++j) crow[j] = 0.0f;
        } else if (beta != 1.0f) {
            for (MKL_INT j = 0; j < n; ++j) crow[j] *= beta;
        }
        for (MKL_INT p = pntrb[r]; p < pntre[r]; ++p) {
            const float a = alpha * val[p];
            const float *brow = b + (size_t)indx[p] * (size_t)ldb;
            for (MKL_INT j = 0; j < n; ++j) crow[j] += a * brow[j];
        }
    }
}

void mkl_scscmm(const char *, const MKL_INT *, const MKL_INT *, const MKL_INT *, const float *,
                const char *, const float *, const MKL_INT *, const MKL_INT *, const MKL_INT *,
                const float *, const MKL_INT *, const float *, float *, const MKL_INT *)
{
    std::abort(); /* aliased by include/types.h:51 but never called */
}

/* job[0]=1: CSC(acsc,ja1,ia1) -> CSR(acsr,ja,ia); zero-based when job[1]=job[2]=0. n x n. */
void mkl_scsrcsc(const MKL_INT *job, const MKL_INT *n_, float *acsr, MKL_INT *ja, MKL_INT *ia,
                 float *acsc, MKL_INT *ja1, MKL_INT *ia1, MKL_INT *info)
{
    const MKL_INT n = *n_;
    if (job[0] != 1 || job[1] != 0 || job[2] != 0) std::abort();
    const MKL_INT nnz = ia1[n];
    for (MKL_INT r = 0; r <= n; ++r) ia[r] = 0;
    for (MKL_INT p = 0; p < nnz; ++p) ia[ja1[p] + 1]++;
    for (MKL_INT r = 0; r < n; ++r) ia[r + 1] += ia[r];
    MKL_INT *fill = (MKL_INT *)std::malloc(sizeof(MKL_INT) * (size_t)(n + 1));
    std::memcpy(fill, ia, sizeof(MKL_INT) * (size_t)(n + 1));
    for (MKL_INT c = 0; c < n; ++c)
        for (MKL_INT p = ia1[c]; p < ia1[c + 1]; ++p) {
            const MKL_INT dst = fill[ja1[p]]++;
            ja[dst] = c;
            if (job[5] != 0) acsr[dst] = acsc[p];
        }
    std::free(fill);
    if (info) *info = 0;
}

/* dense (row-major m x n, lda) -> CSR; job[0]=0, zero-based when job[1]=job[2]=0;
 * job[3]=2 whole matrix; job[4]=nzmax; job[5]: 0 => only ia, >0 => fill all. */
void mkl_sdnscsr(const MKL_INT *job, const MKL_INT *m_, const MKL_INT *n_, float *adns,
                 const MKL_INT *lda_, float *acsr, MKL_INT *ja, MKL_INT *ia, MKL_INT *info)
{
    if (job[0] != 0) std::abort();
    const MKL_INT m = *m_, n = *n_, lda = *lda_;
    const MKL_INT base = job[2] ? 1 : 0;
    const MKL_INT adns_base_is_one = job[1];
    (void)adns_base_is_one;
    MKL_INT pos = 0;
    for (MKL_INT r = 0; r < m; ++r) {
        ia[r] = pos + base;
        for (MKL_INT c = 0; c < n; ++c) {
            const float v = adns[(size_t)r * (size_t)lda + (size_t)c];
            if (v != 0.0f) {
                if (job[5] > 0) {
                    if (pos >= job[4]) {
                        if (info) *info = r + 1;
                        return;
                    }
                    acsr[pos] = v;
                    ja[pos] = c + base;
                }
                ++pos;
            }
        }
    }
    ia[m] = pos + base;
    if (info) *info = 0;
}

void mkl_cspblas_scsrgemv(const char *transa, const MKL_INT *m_, const float *a,
                          const MKL_INT *ia, const MKL_INT *ja, const float *x, float *y)
{
    if (*transa != 'N' && *transa != 'n') std::abort();
    const MKL_INT m = *m_;
#pragma omp parallel for schedule(dynamic, 1024)
    for (MKL_INT r = 0; r < m; ++r) {
        float s = 0.0f;
        for (MKL_INT p = ia[r]; p < ia[r + 1]; ++p) s += a[p] * x[ja[p]];
        y[r] = s;
    }
}

/* B = alpha * op(A); ordering 'C' (column-major) or 'R'; trans 'N' or 'T'. */
void mkl_somatcopy(const char ordering, const char trans, size_t rows, size_t cols,
                   const float alpha, const float *A, size_t lda, float *B, size_t ldb)
{
    const bool colmajor = (ordering == 'C' || ordering == 'c');
    const bool tr = (trans == 'T' || trans == 't');
    /* view A as `outer` vectors of `inner` contiguous elements */
    const size_t outer = colmajor ? cols : rows;
    const size_t inner = colmajor ? rows : cols;
#pragma omp parallel for schedule(static)
    for (long long o = 0; o < (long long)outer; ++o)
        for (size_t i = 0; i < inner; ++i) {
            const float v = alpha * A[(size_t)o * lda + i];
            if (tr)
                B[i * ldb + (size_t)o] = v;
            else
                B[(size_t)o * ldb + i] = v;
        }
}

/* Reference pins MKL to one thread inside an OpenMP region (matUtils.h:352,361).
 * Our mkl_scsrmm uses OpenMP, which is already serial inside a parallel region
 * (nested parallelism is off by default); only the OpenBLAS pool is adjusted. */
int mkl_set_num_threads_local(int nt)
{
    static int saved = 0;
    if (nt > 0) {
        saved = scipy_openblas_get_num_threads64_();
        scipy_openblas_set_num_threads64_(nt);
    } else if (saved > 0) {
        scipy_openblas_set_num_threads64_(saved);
    }
    return 0;
}

void *mkl_malloc(size_t size, int align)
{
    void *p = NULL;
    if (align < (int)sizeof(void *)) align = sizeof(void *);
    if (posix_memalign(&p, (size_t)align, size ? size : 1) != 0) return NULL;
    return p;
}
void mkl_free(void *p) { std::free(p); }

MKL_INT LAPACKE_sgesvd(int, char, char, MKL_INT, MKL_INT, float *, MKL_INT, float *, float *,
                       MKL_INT, float *, MKL_INT, float *)
{
    std::abort(); /* only reachable from diagnostics the driver never calls */
}
} /* extern "C" */
