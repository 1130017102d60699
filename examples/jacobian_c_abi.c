/* jacobian_c_abi.c -- the fused Jacobian from plain C through include/qcdmet_b200.h.
 *
 *   gcc -O2 -I include examples/jacobian_c_abi.c -o /tmp/jacobian_c_abi \
 *       -L qcdmet_b200 -lqcdmet_b200 -Wl,-rpath,$PWD/qcdmet_b200 -lm
 *
 * Builds a small symmetric H0 with a gap, 4-site fragments with the parameter list of
 * src/dmet.py:118-159 (non translation-invariant), calls qcdmet_jacobian once and checks two
 * size-independent properties of the result (block symmetry, zero trace).  Needs a CUDA device. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "qcdmet_b200.h"

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 64, m = 4, F = N / m, no = N / 2;
    const int P = F * m * (m + 1) / 2, Q = F * m * m;
    double *H0 = calloc((size_t)N * N, sizeof(double)), *J = malloc(sizeof(double) * (size_t)Q * P);
    int *h1start = malloc(sizeof(int) * (P + 1)), *h1row = malloc(sizeof(int) * 2 * P), *h1col = malloc(sizeof(int) * 2 * P);
    int *rowStart = malloc(sizeof(int) * (F + 1)), *rowIdx = malloc(sizeof(int) * N);
    /* dimerised chain: alternating hoppings open a gap at half filling */
    for (int i = 0; i + 1 < N; ++i) H0[i * N + i + 1] = H0[(i + 1) * N + i] = (i % 2 == 0) ? -1.0 : -0.6;
    for (int i = 0; i < N; ++i) H0[i * N + i] = 0.05 * sin(1.7 * i);
    int p = 0, nnz = 0;
    h1start[0] = 0;
    for (int f = 0; f < F; ++f) {
        rowStart[f] = f * m;
        for (int r = 0; r < m; ++r) {
            rowIdx[f * m + r] = f * m + r;
            for (int c = r; c < m; ++c) {      /* E_rc + E_cr, elements in row-major order */
                h1row[nnz] = f * m + r; h1col[nnz++] = f * m + c;
                if (c != r) { h1row[nnz] = f * m + c; h1col[nnz++] = f * m + r; }
                h1start[++p] = nnz;
            }
        }
    }
    rowStart[F] = N;
    int rc = qcdmet_jacobian(N, P, no, h1start, h1row, h1col, H0, F, rowStart, NULL, rowIdx, 0, P, J, NULL, NULL);
    if (rc != QCDMET_OK) { fprintf(stderr, "qcdmet_jacobian failed (%d): %s\n", rc, qcdmet_last_error()); return 1; }
    double asym = 0.0, trmax = 0.0;
    for (int col = 0; col < P; ++col) {
        double tr = 0.0;
        for (int f = 0; f < F; ++f)
            for (int a = 0; a < m; ++a) {
                tr += J[(size_t)Q * col + f * m * m + a + m * a];
                for (int b = 0; b < m; ++b)
                    asym = fmax(asym, fabs(J[(size_t)Q * col + f * m * m + a + m * b] - J[(size_t)Q * col + f * m * m + b + m * a]));
            }
        trmax = fmax(trmax, fabs(tr));
    }
    printf("%s: N=%d P=%d Q=%d  block asymmetry %.1e  max |trace| %.1e\n", qcdmet_version(), N, P, Q, asym, trmax);
    return (asym == 0.0 && trmax < 1e-10) ? 0 : 2;
}
