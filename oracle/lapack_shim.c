/* oracle/lapack_shim.c -- TEST INFRASTRUCTURE (not product code).
 *
 * The prebuilt reference `cathy` ELFs shipped under /root/reference link
 * liblapack.so.3 for three symbols only (dcopy_, dgetrf_, dgetrs_), all of
 * them called from the solute-transport module (SRC/iperplane.f:76-77,
 * SRC/dxpay.f:16) which is never entered with TRAFLAG=0.  This shim supplies
 * plain-C versions so the ELF loads without a system LAPACK.
 */
#include <math.h>
#include <stdlib.h>

void dcopy_(const int *n, const double *x, const int *incx, double *y, const int *incy)
{
    int ix = (*incx < 0) ? (1 - *n) * (*incx) : 0;
    int iy = (*incy < 0) ? (1 - *n) * (*incy) : 0;
    for (int i = 0; i < *n; ++i, ix += *incx, iy += *incy) y[iy] = x[ix];
}

/* column-major LU with partial pivoting (unblocked) */
void dgetrf_(const int *m, const int *n, double *a, const int *lda, int *ipiv, int *info)
{
    int M = *m, N = *n, L = *lda, mn = M < N ? M : N;
    *info = 0;
    for (int j = 0; j < mn; ++j) {
        int p = j;
        double big = fabs(a[j + j * L]);
        for (int i = j + 1; i < M; ++i)
            if (fabs(a[i + j * L]) > big) { big = fabs(a[i + j * L]); p = i; }
        ipiv[j] = p + 1;
        if (big == 0.0) { if (*info == 0) *info = j + 1; continue; }
        if (p != j)
            for (int k = 0; k < N; ++k) { double t = a[j + k * L]; a[j + k * L] = a[p + k * L]; a[p + k * L] = t; }
        for (int i = j + 1; i < M; ++i) a[i + j * L] /= a[j + j * L];
        for (int k = j + 1; k < N; ++k)
            for (int i = j + 1; i < M; ++i) a[i + k * L] -= a[i + j * L] * a[j + k * L];
    }
}

void dgetrs_(const char *trans, const int *n, const int *nrhs, const double *a, const int *lda,
             const int *ipiv, double *b, const int *ldb, int *info, int trans_len)
{
    int N = *n, L = *lda, LB = *ldb;
    (void)trans_len;
    *info = 0;
    if (*trans != 'N' && *trans != 'n') { *info = -1; return; }
    for (int r = 0; r < *nrhs; ++r) {
        double *x = b + (size_t)r * LB;
        for (int i = 0; i < N; ++i) { int p = ipiv[i] - 1; if (p != i) { double t = x[i]; x[i] = x[p]; x[p] = t; } }
        for (int i = 0; i < N; ++i) for (int k = 0; k < i; ++k) x[i] -= a[i + k * L] * x[k];
        for (int i = N - 1; i >= 0; --i) { for (int k = i + 1; k < N; ++k) x[i] -= a[i + k * L] * x[k]; x[i] /= a[i + i * L]; }
    }
}
