/*
 * TEST INFRASTRUCTURE — not product code.
 *
 * Minimal Fortran-ABI `dgemm_` so the UNMODIFIED reference DCI sources
 * (/root/reference/dci_code/src/{dci,util,py_dci}.c) link without a system BLAS.
 *
 * The reference calls BLAS at exactly one place: util.c:34-43 `matmul`, always with
 * TRANSA='T', TRANSB='N', ALPHA=1, BETA=0 (declared in dci_code/include/util.h:35-42),
 * to project data/queries onto the m*L random directions (dci.c:220,234,799).
 * The projections only steer WHICH candidates DCI visits; the distances it reports come
 * from util.c:62-69 compute_dist.  The reference leaves the BLAS vendor unpinned
 * (Makefile:33-35,96-109: netlib|atlas|openblas|mkl), so any correct dgemm is conforming.
 *
 * Only the ('T','N') case is implemented; anything else aborts loudly.
 * Column-major: A is K x M (lda), B is K x N (ldb), C is M x N (ldc);  C = alpha*A^T*B + beta*C.
 */
#include <stdio.h>
#include <stdlib.h>

void dgemm_(const char *transa, const char *transb, const int *pm, const int *pn, const int *pk,
            const double *palpha, const double *A, const int *plda, const double *B, const int *pldb,
            const double *pbeta, double *C, const int *pldc)
{
    const int M = *pm, N = *pn, K = *pk;
    const long lda = *plda, ldb = *pldb, ldc = *pldc;
    const double alpha = *palpha, beta = *pbeta;
    if ((*transa != 'T' && *transa != 't') || (*transb != 'N' && *transb != 'n')) {
        fprintf(stderr, "oracle/dgemm_shim: only dgemm('T','N') is implemented (got '%c','%c')\n", *transa, *transb);
        abort();
    }
    long n;
#pragma omp parallel for schedule(static)
    for (n = 0; n < N; n++) {
        const double *b = B + n * ldb;
        int m;
        for (m = 0; m < M; m++) {
            const double *a = A + m * lda;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = 0;
            for (; k + 4 <= K; k += 4) {
                s0 += a[k] * b[k];
                s1 += a[k + 1] * b[k + 1];
                s2 += a[k + 2] * b[k + 2];
                s3 += a[k + 3] * b[k + 3];
            }
            for (; k < K; k++) s0 += a[k] * b[k];
            const double s = (s0 + s1) + (s2 + s3);
            double *c = C + m + n * ldc;
            *c = (beta == 0.0) ? alpha * s : alpha * s + beta * (*c);
        }
    }
}
