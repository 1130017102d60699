/* TEST INFRASTRUCTURE ONLY (oracle/): not part of the product path.
 *
 * The MKL embedded in torch's libtorch_cpu.so exports dsyevd_ but not dsyev_,
 * which /root/reference/lib/rhf_response.cc:26,50 calls.  This shim provides
 * dsyev_ on top of dsyevd_ (divide & conquer; same eigen-decomposition up to
 * rounding and eigenvector sign), so the UNMODIFIED reference source links
 * against that MKL ("MKL-built libqcdmet" of BASELINE.json.north_star).
 */
#include <stdlib.h>

void dsyevd_(char *jobz, char *uplo, const int *n, double *a, const int *lda,
             double *w, double *work, int *lwork, int *iwork, int *liwork,
             int *info);

void dsyev_(char *jobz, char *uplo, const int *n, double *a, const int *lda,
            double *w, double *work_unused, int *lwork_unused, int *info)
{
    (void)work_unused; (void)lwork_unused;
    double wq = 0.0; int iwq = 0, m1 = -1;
    dsyevd_(jobz, uplo, n, a, lda, w, &wq, &m1, &iwq, &m1, info);
    if (*info != 0) return;
    int lwork = (int)wq, liwork = iwq;
    double *work = (double *)malloc(sizeof(double) * (size_t)(lwork > 1 ? lwork : 1));
    int *iwork = (int *)malloc(sizeof(int) * (size_t)(liwork > 1 ? liwork : 1));
    dsyevd_(jobz, uplo, n, a, lda, w, work, &lwork, iwork, &liwork, info);
    free(work); free(iwork);
}
