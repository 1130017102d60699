// kernels_bath.cu -- Schmidt baths of ALL fragments in one call, from the mean-field orbitals on the device
// (replaces the per-fragment numpy eigh of src/qcdmethelper.py:124-168).
//
// The reference diagonalises the (N-m) x (N-m) environment block of D = 2 C_o C_o^T for every fragment and
// keeps the m eigenvectors whose occupation is furthest from 0 and 2.  P = D/2 is a projector, so with I the
// impurity orbitals and E the environment
//     P_EE P_EI + P_EI P_II = P_EI      (EI block of P^2 = P)
// and for an eigenpair P_II u_k = mu_k u_k the vector P_EI u_k is an eigenvector of P_EE with eigenvalue
// 1 - mu_k and squared norm mu_k (1 - mu_k).  Every environment natural orbital with an occupation strictly
// between 0 and 2 is of that form, hence:
//     bath_k = P_EI u_k / |P_EI u_k|,   occupation 2 (1 - mu_k),   k ranked by min(occ, 2 - occ), largest first
// (the ordering rule of src/qcdmethelper.py:139-141), from the eigenvectors of an m x m matrix instead of an
// (N-m) x (N-m) one.  Exact for an idempotent D/2, which is what construct1RDM_loc returns (:114-122).
//
// One CTA per fragment: Gram matrix P_II = PI PI^T of the impurity rows PI = C_o[I, :], parallel two-sided Jacobi
// (round-robin pair ordering, rotations of disjoint pairs applied together) in shared memory, ranking, and
// Y_k = u_k^T PI; then ONE library DGEMM  B = Y C_o^T  for all fragments (plain GEMM: cuBLAS), and a
// normalise / mask kernel (impurity entries of a bath orbital are zero).
#include "common.cuh"
#include <algorithm>

namespace qcdmet {

namespace {

constexpr int BATH_THREADS = 256;
constexpr int BATH_MAX_SWEEPS = 60;

// shared memory: A [mp][mp+1], V [mp][mp+1], rotation table (c, s, p, q) per pair, ranking
__global__ void __launch_bounds__(BATH_THREADS) bath_eigen_kernel(
    const double *__restrict__ CoT, int ldo, int no, const int *__restrict__ impStart, const int *__restrict__ impIdx,
    const int *__restrict__ numReq, double threshold, double *__restrict__ Y /* [S][ldo] */,
    double *__restrict__ occOut /* [S] */, int *__restrict__ numKept /* [F] */, int *__restrict__ sweepsOut /* [F] */)
{
    extern __shared__ __align__(16) unsigned char bsm[];
    const int f = blockIdx.x;
    const int r0 = impStart[f], m = impStart[f + 1] - r0;
    if (m == 0) { if (threadIdx.x == 0) { numKept[f] = 0; sweepsOut[f] = 0; } return; }
    const int mp = (m + 1) & ~1, ld = mp + 1, npairs = mp / 2;
    QC_DEV_ASSERT(m <= QCDMET_BATH_MAX_IMP);
    double *A = reinterpret_cast<double *>(bsm);
    double *V = A + mp * ld;
    double *cs = V + mp * ld;                        // [npairs][2]
    int *pq = reinterpret_cast<int *>(cs + 2 * npairs);   // [npairs][2]
    int *order = pq + 2 * npairs;                    // [mp]
    int *flag = order + mp;                          // rotations applied in this sweep
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = BATH_THREADS / 32;

    // ---- Gram matrix of the impurity rows: one warp per (i <= j) pair
    for (int pr = warp; pr < m * (m + 1) / 2; pr += nwarps) {
        int i = 0, rem = pr;
        while (rem >= m - i) { rem -= m - i; ++i; }
        const int j = i + rem;
        const double *pa = CoT + (size_t)impIdx[r0 + i] * ldo, *pb = CoT + (size_t)impIdx[r0 + j] * ldo;
        double s = 0.0;
        for (int o = lane; o < no; o += 32) s += pa[o] * pb[o];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) { A[i * ld + j] = s; A[j * ld + i] = s; }
    }
    for (int idx = tid; idx < mp * mp; idx += BATH_THREADS) {
        const int i = idx / mp, j = idx % mp;
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
        if (i >= m || j >= m) A[i * ld + j] = 0.0;   // padding index of an odd m: never rotated
    }
    __syncthreads();

    // ---- parallel cyclic Jacobi: mp - 1 steps per sweep, npairs disjoint rotations per step
    int sweeps = 0;
    for (; sweeps < BATH_MAX_SWEEPS; ++sweeps) {
        if (tid == 0) *flag = 0;
        __syncthreads();
        for (int step = 0; step < mp - 1; ++step) {
            if (tid < npairs) {
                // circle method: player mp-1 stays, the others rotate
                int p = tid == 0 ? mp - 1 : (step + tid) % (mp - 1);
                int q = (step - tid + (mp - 1)) % (mp - 1);
                if (p > q) { const int t = p; p = q; q = t; }
                QC_DEV_ASSERT(p >= 0 && q < mp && p != q);
                double c = 1.0, s = 0.0;
                if (q < m) {
                    const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
                    if (fabs(apq) > 1e-16 * sqrt(fabs(app) * fabs(aqq)) && fabs(apq) > 1e-290) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                        *flag = 1;
                    }
                }
                cs[2 * tid] = c; cs[2 * tid + 1] = s;
                pq[2 * tid] = p; pq[2 * tid + 1] = q;
            }
            __syncthreads();
            // columns: (x_p, x_q) <- (c x_p - s x_q, s x_p + c x_q) in every row of A and V
            for (int idx = tid; idx < npairs * mp; idx += BATH_THREADS) {
                const int pr = idx / mp, r = idx % mp;
                const double c = cs[2 * pr], s = cs[2 * pr + 1];
                if (s == 0.0) continue;
                const int p = pq[2 * pr], q = pq[2 * pr + 1];
                const double ap = A[r * ld + p], aq = A[r * ld + q];
                A[r * ld + p] = c * ap - s * aq;
                A[r * ld + q] = s * ap + c * aq;
                const double vp = V[r * ld + p], vq = V[r * ld + q];
                V[r * ld + p] = c * vp - s * vq;
                V[r * ld + q] = s * vp + c * vq;
            }
            __syncthreads();
            // rows of A
            for (int idx = tid; idx < npairs * mp; idx += BATH_THREADS) {
                const int pr = idx / mp, r = idx % mp;
                const double c = cs[2 * pr], s = cs[2 * pr + 1];
                if (s == 0.0) continue;
                const int p = pq[2 * pr], q = pq[2 * pr + 1];
                const double ap = A[p * ld + r], aq = A[q * ld + r];
                A[p * ld + r] = c * ap - s * aq;
                A[q * ld + r] = s * ap + c * aq;
            }
            __syncthreads();
        }
        if (*flag == 0) break;
        __syncthreads();
    }

    // ---- ranking: depth = min(occ, 2 - occ) of the environment occupation occ = 2 (1 - mu), largest first;
    //      ties: smaller occupation first (the order a stable sort of ascending eigenvalues gives)
    if (tid == 0) {
        for (int k = 0; k < m; ++k) order[k] = k;
        for (int a = 1; a < m; ++a) {                 // insertion sort, m <= 64
            const int ka = order[a];
            const double oa = 2.0 * (1.0 - A[ka * ld + ka]), da = fmin(oa, 2.0 - oa);
            int b = a - 1;
            while (b >= 0) {
                const int kb = order[b];
                const double ob = 2.0 * (1.0 - A[kb * ld + kb]), db = fmin(ob, 2.0 - ob);
                if (db > da || (db == da && ob <= oa)) break;
                order[b + 1] = kb;
                --b;
            }
            order[b + 1] = ka;
        }
        int keep = 0;
        for (int k = 0; k < m; ++k) {
            const double o = 2.0 * (1.0 - A[order[k] * ld + order[k]]);
            if (fmin(o, 2.0 - o) > threshold) ++keep;
        }
        const int req = numReq ? numReq[f] : m;
        numKept[f] = min(keep, req);
        sweepsOut[f] = sweeps;
    }
    __syncthreads();
    for (int k = tid; k < m; k += BATH_THREADS) occOut[r0 + k] = 2.0 * (1.0 - A[order[k] * ld + order[k]]);
    // ---- Y_k[o] = sum_i u_k[i] PI[i][o]  (rank order); zero rows for orbitals that are not kept
    const int kept = numKept[f];
    for (int o = tid; o < ldo; o += BATH_THREADS) {
        for (int k0 = 0; k0 < m; k0 += 8) {
            double acc[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) acc[kk] = 0.0;
            if (o < no)
                for (int i = 0; i < m; ++i) {
                    const double x = CoT[(size_t)impIdx[r0 + i] * ldo + o];
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        if (k0 + kk < kept) acc[kk] = fma(V[i * ld + order[k0 + kk]], x, acc[kk]);
                }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
                if (k0 + kk < m) Y[(size_t)(r0 + k0 + kk) * ldo + o] = acc[kk];
        }
    }
}

// B[k][:] (one CTA per bath row): zero on the fragment's own impurity orbitals, unit norm elsewhere
__global__ void bath_normalise_kernel(double *__restrict__ B, int N, const int *__restrict__ impStart,
                                      const int *__restrict__ impIdx, const int *__restrict__ rowFrag,
                                      const int *__restrict__ numKept)
{
    __shared__ double red[32];
    const int k = blockIdx.x, f = rowFrag[k];
    const int r0 = impStart[f], m = impStart[f + 1] - r0;
    double *row = B + (size_t)k * N;
    if (k - r0 >= numKept[f]) {                       // not a bath orbital: zero row
        for (int x = threadIdx.x; x < N; x += blockDim.x) row[x] = 0.0;
        return;
    }
    if (threadIdx.x < m) row[impIdx[r0 + threadIdx.x]] = 0.0;
    for (int i = threadIdx.x + blockDim.x; i < m; i += blockDim.x) row[impIdx[r0 + i]] = 0.0;
    __syncthreads();
    double s = 0.0;
    for (int x = threadIdx.x; x < N; x += blockDim.x) s += row[x] * row[x];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    const double scale = red[0] > 0.0 ? rsqrt(red[0]) : 0.0;
    for (int x = threadIdx.x; x < N; x += blockDim.x) row[x] *= scale;
}

}  // namespace

size_t bath_smem_bytes(int maxM)
{
    const int mp = (maxM + 1) & ~1, ld = mp + 1;
    return (size_t)2 * mp * ld * 8 + (size_t)mp * 8 /* cs */ + (size_t)mp * 4 /* pq */ + (size_t)mp * 4 /* order */ + 16;
}

// Y, occ, numKept, sweeps, B are device buffers; the index arrays too.  B = Y CoT^T through cuBLAS on `st`.
int launch_construct_baths(cublasHandle_t blas, const double *CoT, int ldo, int N, int no, int F, int S, int maxM,
                           const int *d_impStart, const int *d_impIdx, const int *d_rowFrag, const int *d_numReq,
                           double threshold, double *d_Y, double *d_occ, int *d_numKept, int *d_sweeps, double *d_B,
                           cudaStream_t st)
{
    if (F == 0 || S == 0) return QCDMET_OK;
    const size_t smem = bath_smem_bytes(maxM);
    QC_CUDA(cudaFuncSetAttribute(bath_eigen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bath_eigen_kernel<<<F, BATH_THREADS, smem, st>>>(CoT, ldo, no, d_impStart, d_impIdx, d_numReq, threshold, d_Y, d_occ,
                                                     d_numKept, d_sweeps);
    QC_CUDA(cudaGetLastError());
    // B (row-major [S][N]) = Y ([S][ldo]) CoT^T:  column-major  B_cm (N x S) = CoT_cm^T (N x no) Y_cm (no x S)
    const double one = 1.0, zero = 0.0;
    cublasSetStream(blas, st);
    cublasStatus_t cs = cublasDgemm(blas, CUBLAS_OP_T, CUBLAS_OP_N, N, S, no, &one, CoT, ldo, d_Y, ldo, &zero, d_B, N);
    QC_CHECK(cs == CUBLAS_STATUS_SUCCESS, QCDMET_E_CUDA, "cublasDgemm (bath orbitals) failed: %d", (int)cs);
    bath_normalise_kernel<<<S, 256, 0, st>>>(d_B, N, d_impStart, d_impIdx, d_rowFrag, d_numKept);
    QC_CUDA(cudaGetLastError());
    return QCDMET_OK;
}

}  // namespace qcdmet
