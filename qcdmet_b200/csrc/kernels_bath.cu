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
    const double tol = 4.0 * 2.220446049250313e-16 * sqrt((double)m);
    double scale = 0.0;
    for (int i = 0; i < m; ++i) scale = fmax(scale, fabs(A[i * ld + i]));
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
                    // rotate until |a_pq| <= tol max_i a_ii: absolute accuracy of a few ulps of the largest eigenvalue
                    // (<= 1 for a projector block).  A criterion relative to sqrt(a_pp a_qq) never ends in the null space
                    // of P_II (more impurity orbitals than occupied orbitals), where all three are rounding noise.
                    if (fabs(apq) > tol * scale) {
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


// ------------------------------------------------------------------------------------------
// Embedded Hamiltonians of all fragments for an on-site (Hubbard) interaction, from the device baths
// (src/localintegrals_hubbard.py:88-116: dmet_oei, dmet_fock with loc_rhf_fock_bis(core 1-RDM), dmet_tei, whose
// O(n^4 N) Python loop per fragment is the second per-iteration wall after the bath eigh).  With c = [unit vectors of
// the impurity orbitals | bath orbitals] (bath orbitals vanish on the impurity):
//   oei  = c^T h c                     imp-imp: h[i,j];  imp-bath: (B h)[k, i];  bath-bath: B (h B^T)
//   fock = oei + (U/2) c^T diag(core) c, core[x] = D[x,x] - sum_k occ_k B[k,x]^2 on the environment, 0 on the impurity:
//          only the bath-bath block changes
//   tei[a,b,c,d] = U sum_x c_xa c_xb c_xc c_xd: U on the impurity diagonal, U sum_x B_k B_l B_p B_q on the bath block,
//          zero wherever impurity and bath indices mix
// HB = B h comes from one library DGEMM for all fragments; this kernel does the per-fragment reductions.
// ------------------------------------------------------------------------------------------
namespace {

constexpr int EMB_THREADS = 256;
constexpr int EMB_XCH = 128;            // lattice orbitals per staged chunk

__global__ void rdm_diag_kernel(const double *__restrict__ CoT, int ldo, int no, int N, double *__restrict__ ddiag)
{
    const int x = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (x >= N) return;
    const double *row = CoT + (size_t)x * ldo;
    double s = 0.0;
    for (int o = lane; o < no; o += 32) s += row[o] * row[o];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) ddiag[x] = 2.0 * s;
}

// one CTA per fragment; outputs are packed per fragment with capacity ncap = m + (requested bath orbitals)
__global__ void __launch_bounds__(EMB_THREADS) embed_hubbard_kernel(
    const double *__restrict__ H /* N x N symmetric */, const double *__restrict__ B /* [S][N] */,
    const double *__restrict__ HB /* [S][N] = B h */, const double *__restrict__ ddiag, const double *__restrict__ occ,
    int N, double U, const int *__restrict__ impStart, const int *__restrict__ impIdx, const int *__restrict__ numKept,
    const int *__restrict__ cap /* [F] */, const long long *__restrict__ off2 /* [F] */, const long long *__restrict__ off4 /* [F], -1 = no tei */,
    int nbMax, double *__restrict__ oei, double *__restrict__ fock, double *__restrict__ tei, double *__restrict__ coreTrace)
{
    extern __shared__ __align__(16) unsigned char esm[];  // bath rows of the chunk and their images under h: [nbMax][XCH + 1] each
    __shared__ double sCore[EMB_XCH];
    __shared__ double red[EMB_THREADS / 32];
    typedef double Row[EMB_XCH + 1];
    Row *sB = reinterpret_cast<Row *>(esm);
    Row *sHB = sB + nbMax;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int r0 = impStart[f], m = impStart[f + 1] - r0, nb = numKept[f], n = cap[f];
    QC_DEV_ASSERT(nb <= nbMax && m + nb <= n);
    double *o1 = oei + off2[f], *f1 = fock + off2[f];
    for (int i = tid; i < n * n; i += EMB_THREADS) { o1[i] = 0.0; f1[i] = 0.0; }
    const bool wantTei = off4[f] >= 0;
    double *t4 = wantTei ? tei + off4[f] : nullptr;
    if (wantTei)
        for (long long i = tid; i < (long long)n * n * n * n; i += EMB_THREADS) t4[i] = 0.0;
    __syncthreads();
    // impurity-impurity and impurity-bath blocks (row-major n x n)
    for (int idx = tid; idx < m * m; idx += EMB_THREADS) {
        const int a = idx / m, b = idx % m;
        const double v = H[(size_t)impIdx[r0 + a] * N + impIdx[r0 + b]];
        o1[a * n + b] = v; f1[a * n + b] = v;
    }
    for (int idx = tid; idx < m * nb; idx += EMB_THREADS) {
        const int a = idx / nb, k = idx % nb;
        const double v = HB[(size_t)(r0 + k) * N + impIdx[r0 + a]];
        o1[a * n + (m + k)] = v; o1[(m + k) * n + a] = v;
        f1[a * n + (m + k)] = v; f1[(m + k) * n + a] = v;
    }
    if (wantTei)
        for (int a = tid; a < m; a += EMB_THREADS) t4[(((long long)a * n + a) * n + a) * n + a] = U;
    // bath-bath blocks: reductions over the lattice in chunks
    const int nbb = nb * nb;
    const int per2 = (nbb + EMB_THREADS - 1) / EMB_THREADS;           // (k, l) pairs per thread: 1 for nb <= 16
    const long long n4 = wantTei ? (long long)nbb * nbb : 0;
    const int per4 = (int)((n4 + EMB_THREADS - 1) / EMB_THREADS);     // <= 16 for nb <= 8
    double accO[16], accF[16], accT[16];
    for (int i = 0; i < 16; ++i) accO[i] = accF[i] = accT[i] = 0.0;
    double ctr = 0.0;
    for (int x0 = 0; x0 < N; x0 += EMB_XCH) {
        const int nx = min(EMB_XCH, N - x0);
        for (int idx = tid; idx < nb * EMB_XCH; idx += EMB_THREADS) {
            const int k = idx / EMB_XCH, xx = idx % EMB_XCH;
            sB[k][xx] = xx < nx ? B[(size_t)(r0 + k) * N + x0 + xx] : 0.0;
            sHB[k][xx] = xx < nx ? HB[(size_t)(r0 + k) * N + x0 + xx] : 0.0;
        }
        __syncthreads();
        // core occupation of the chunk's orbitals: D[x,x] - sum_k occ_k B[k,x]^2 (bath rows are zero on the impurity, where
        // the core density vanishes: mask those orbitals explicitly)
        for (int xx = tid; xx < EMB_XCH; xx += EMB_THREADS) {
            double c = 0.0;
            if (xx < nx) {
                c = ddiag[x0 + xx];
                for (int k = 0; k < nb; ++k) c -= occ[r0 + k] * sB[k][xx] * sB[k][xx];
                for (int a = 0; a < m; ++a)
                    if (impIdx[r0 + a] == x0 + xx) c = 0.0;
            }
            sCore[xx] = c;
            ctr += c;
        }
        __syncthreads();
        for (int q = 0; q < per2 && q < 16; ++q) {
            const int pr = tid + q * EMB_THREADS;
            if (pr >= nbb) break;
            const int k = pr / nb, l = pr % nb;
            double so = 0.0, sf = 0.0;
            for (int xx = 0; xx < EMB_XCH; ++xx) {
                so = fma(sB[k][xx], sHB[l][xx], so);
                sf = fma(sB[k][xx] * sB[l][xx], sCore[xx], sf);
            }
            accO[q] += so; accF[q] += sf;
        }
        for (int q = 0; q < per4 && q < 16; ++q) {
            const long long id = tid + (long long)q * EMB_THREADS;
            if (id >= n4) break;
            const int kl = (int)(id / nbb), pq = (int)(id % nbb);
            const int k = kl / nb, l = kl % nb, p = pq / nb, qq = pq % nb;
            double st = 0.0;
            for (int xx = 0; xx < EMB_XCH; ++xx) st = fma(sB[k][xx] * sB[l][xx], sB[p][xx] * sB[qq][xx], st);
            accT[q] += st;
        }
        __syncthreads();
    }
    for (int q = 0; q < per2 && q < 16; ++q) {
        const int pr = tid + q * EMB_THREADS;
        if (pr >= nbb) break;
        const int k = pr / nb, l = pr % nb;
        o1[(m + k) * n + (m + l)] = accO[q];
        f1[(m + k) * n + (m + l)] = accO[q] + 0.5 * U * accF[q];
    }
    for (int q = 0; q < per4 && q < 16; ++q) {
        const long long id = tid + (long long)q * EMB_THREADS;
        if (id >= n4) break;
        const int kl = (int)(id / nbb), pq = (int)(id % nbb);
        const int k = kl / nb, l = kl % nb, p = pq / nb, qq = pq % nb;
        t4[((((long long)(m + k)) * n + (m + l)) * n + (m + p)) * n + (m + qq)] = U * accT[q];
    }
    // trace of the core 1-RDM (electrons outside the embedded problem)
    for (int off = 16; off > 0; off >>= 1) ctr += __shfl_xor_sync(0xffffffffu, ctr, off);
    if ((tid & 31) == 0) red[tid >> 5] = ctr;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < EMB_THREADS / 32; ++w) t += red[w];
        coreTrace[f] = t;
    }
}

}  // namespace

// d_B, d_occ, d_numKept: outputs of launch_construct_baths (device).  d_H: hopping matrix on the device.  d_HB: scratch [S][N].
int launch_embed_hubbard(cublasHandle_t blas, const double *CoT, int ldo, int N, int no, int F, int S, const double *d_H, double U,
                         const int *d_impStart, const int *d_impIdx, const int *d_numKept, const int *d_cap,
                         const long long *d_off2, const long long *d_off4, int nbMax, const double *d_B, const double *d_occ, double *d_HB,
                         double *d_ddiag, double *d_oei, double *d_fock, double *d_tei, double *d_coreTrace, cudaStream_t st)
{
    if (F == 0 || S == 0) return QCDMET_OK;
    rdm_diag_kernel<<<(N + 7) / 8, 256, 0, st>>>(CoT, ldo, no, N, d_ddiag);
    QC_CUDA(cudaGetLastError());
    // HB (row-major [S][N]) = B h:  column-major  HB_cm (N x S) = h (N x N, symmetric) B_cm (N x S)
    const double one = 1.0, zero = 0.0;
    cublasSetStream(blas, st);
    cublasStatus_t cs = cublasDgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, N, S, N, &one, d_H, N, d_B, N, &zero, d_HB, N);
    QC_CHECK(cs == CUBLAS_STATUS_SUCCESS, QCDMET_E_CUDA, "cublasDgemm (h times bath orbitals) failed: %d", (int)cs);
    const size_t smem = (size_t)2 * std::max(nbMax, 1) * (EMB_XCH + 1) * sizeof(double);
    QC_CUDA(cudaFuncSetAttribute(embed_hubbard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_hubbard_kernel<<<F, EMB_THREADS, smem, st>>>(d_H, d_B, d_HB, d_ddiag, d_occ, N, U, d_impStart, d_impIdx, d_numKept, d_cap,
                                                      d_off2, d_off4, nbMax, d_oei, d_fock, d_tei, d_coreTrace);
    QC_CUDA(cudaGetLastError());
    return QCDMET_OK;
}

}  // namespace qcdmet
