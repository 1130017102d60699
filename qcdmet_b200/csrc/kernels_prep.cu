// kernels_prep.cu -- operand panels built once per (C, e): site-major orbital panels,
// fragment-row gathers, mean-field 1-RDM blocks.
#include "common.cuh"

namespace qcdmet {

// C is column-major N x N (eigenvector k in column k, cuSOLVER/LAPACK layout).
// CoT[i][o] = C[i + N*o], CvT[i][v] = C[i + N*(no+v)]; classic 32x32 smem-tile transpose so
// both the read (i contiguous) and the write (orbital contiguous) are coalesced.
__global__ void split_transpose_kernel(const double *__restrict__ C, int N, int no,
                                       double *__restrict__ CoT, int ldo,
                                       double *__restrict__ CvT, int ldv)
{
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int dk = threadIdx.y; dk < 32; dk += blockDim.y) {
        int i = i0 + threadIdx.x, k = k0 + dk;
        tile[dk][threadIdx.x] = (i < N && k < N) ? C[(size_t)i + (size_t)N * k] : 0.0;
    }
    __syncthreads();
    for (int di = threadIdx.y; di < 32; di += blockDim.y) {
        int i = i0 + di, k = k0 + threadIdx.x;
        if (i < N && k < N) {
            double x = tile[threadIdx.x][di];
            if (k < no) CoT[(size_t)i * ldo + k] = x;
            else        CvT[(size_t)i * ldv + (k - no)] = x;
        }
    }
}

int launch_split_transpose(const double *C, int N, int no, double *CoT, int ldo, double *CvT,
                           int ldv, cudaStream_t st)
{
    QC_CUDA(cudaMemsetAsync(CoT, 0, sizeof(double) * (size_t)N * ldo, st));
    QC_CUDA(cudaMemsetAsync(CvT, 0, sizeof(double) * (size_t)N * ldv, st));
    dim3 grid((N + 31) / 32, (N + 31) / 32), block(32, 8);
    split_transpose_kernel<<<grid, block, 0, st>>>(C, N, no, CoT, ldo, CvT, ldv);
    QC_CUDA(cudaGetLastError());
    return QCDMET_OK;
}

// dst[a][:] = src[rowIdx[a]][:]   (impurity-only fragment rows, src/dmet.py:509)
__global__ void gather_rows_kernel(const double *__restrict__ src, int ld,
                                   const int *__restrict__ rowIdx, double *__restrict__ dst)
{
    const int a = blockIdx.x;
    const double2 *s = reinterpret_cast<const double2 *>(src + (size_t)rowIdx[a] * ld);
    double2 *d = reinterpret_cast<double2 *>(dst + (size_t)a * ld);
    for (int k = threadIdx.x; k < ld / 2; k += blockDim.x) d[k] = s[k];
}

int launch_gather_rows(const double *src, int ld, const int *rowIdx, int S, double *dst,
                       cudaStream_t st)
{
    if (S == 0) return QCDMET_OK;
    gather_rows_kernel<<<S, 128, 0, st>>>(src, ld, rowIdx, dst);
    QC_CUDA(cudaGetLastError());
    return QCDMET_OK;
}

// D_blocks[q(f;a,b)] = 2 * sum_o PO[a][o] PO[b][o]  -- the fragment blocks of the mean-field
// 1-RDM (model part of src/dmet.py:416-421).  One warp per entry, shuffle reduction.
__global__ void dblocks_kernel(JacView v, double *__restrict__ D)
{
    const long long q = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (q >= v.Q) return;
    const int lane = threadIdx.x & 31;
    int lo = 0, hi = v.F;   // fragment with qoff[f] <= q < qoff[f+1]
    while (hi - lo > 1) { int mid = (lo + hi) / 2; if (v.qoff[mid] <= q) lo = mid; else hi = mid; }
    const int f = lo, rs = v.rowStart[f], rf = v.rowStart[f + 1] - rs;
    const long long loc = q - v.qoff[f];
    const int a = rs + (int)(loc % rf), b = rs + (int)(loc / rf);
    const double *pa = v.PO + (size_t)a * v.ldo, *pb = v.PO + (size_t)b * v.ldo;
    double s = 0.0;
    for (int o = lane; o < v.no; o += 32) s += pa[o] * pb[o];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) D[q] = 2.0 * s;
}

int launch_dblocks(const JacView &v, double *D, cudaStream_t st)
{
    if (v.Q == 0) return QCDMET_OK;
    const int wpb = 8;
    long long blocks = (v.Q + wpb - 1) / wpb;
    dblocks_kernel<<<(unsigned)blocks, wpb * 32, 0, st>>>(v, D);
    QC_CUDA(cudaGetLastError());
    return QCDMET_OK;
}

// W[v][o] = 1 / (e_o - e_v)   (lib/rhf_response.cc:66-71 with the sign folded in), zero in the padding
__global__ void build_w_kernel(const double *__restrict__ e, int no, int nv, int ldo, double *__restrict__ W)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nv * ldo) return;
    const int v = (int)(idx / ldo), o = (int)(idx % ldo);
    W[idx] = (o < no) ? 1.0 / (e[o] - e[no + v]) : 0.0;
}

int launch_build_w(const double *e, int no, int nv, int ldo, double *W, cudaStream_t st)
{
    const size_t total = (size_t)nv * ldo;
    build_w_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e, no, nv, ldo, W);
    QC_CUDA(cudaGetLastError());
    return QCDMET_OK;
}

// Full 1-RDM D[i + N*j] = 2 sum_o CoT[i][o] CoT[j][o]   (compat: lib/rhf_response.cc:57-64)
__global__ void one_rdm_kernel(const double *__restrict__ CoT, int ldo, int N, int no,
                               double *__restrict__ D)
{
    __shared__ double sa[16][17], sb[16][17];
    const int i = blockIdx.x * 16 + threadIdx.x, j = blockIdx.y * 16 + threadIdx.y;
    double acc = 0.0;
    for (int o0 = 0; o0 < no; o0 += 16) {
        int ia = blockIdx.x * 16 + threadIdx.y, jb = blockIdx.y * 16 + threadIdx.y;
        int o = o0 + threadIdx.x;
        sa[threadIdx.y][threadIdx.x] = (ia < N && o < no) ? CoT[(size_t)ia * ldo + o] : 0.0;
        sb[threadIdx.y][threadIdx.x] = (jb < N && o < no) ? CoT[(size_t)jb * ldo + o] : 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += sa[threadIdx.x][k] * sb[threadIdx.y][k];
        __syncthreads();
    }
    if (i < N && j < N) D[(size_t)i + (size_t)N * j] = 2.0 * acc;
}

int launch_one_rdm(const double *CoT, int ldo, int N, int no, double *D, cudaStream_t st)
{
    dim3 grid((N + 15) / 16, (N + 15) / 16), block(16, 16);
    one_rdm_kernel<<<grid, block, 0, st>>>(CoT, ldo, N, no, D);
    QC_CUDA(cudaGetLastError());
    return QCDMET_OK;
}

}  // namespace qcdmet
