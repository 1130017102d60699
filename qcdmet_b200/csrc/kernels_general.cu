// kernels_general.cu -- general FP64 fragment-block response kernel.
//
// Correct for arbitrary CSR perturbations (nnz_p >= 1, translation-invariant lists included),
// arbitrary fragment row lists and dense projector rows (fitImpBath).  Used for the shapes neither the
// fused DMMA kernel (kernels_dmma.cu) nor the factorised path (panel mode + kernels_contract.cu) covers
// -- perturbations of more than 64 elements, panels that do not fit in device memory -- and as the
// on-device cross-check of both at sizes the CPU oracle cannot reach.  Plain FP64-FMA-pipe kernel (on B200 the DFMA and DMMA rates are the
// same pipe: 33.7 vs 36.9 TFLOP/s measured), shared-memory tiled, 64 x 64 output tile, 4 x 4 per thread.
//
// Per column p (lib/rhf_response.cc:82-108 restricted to the fragment rows):
//   Z_p[v,o] = (sum_{(r,c) in H1_p} Cv[r,v] Co[c,o]) * W[v,o],   W[v,o] = 1/(e_o - e_v) precomputed once
//   Y_p[a,o] = sum_v PV[a,v] Z_p[v,o]
//   G_p[a,b] = sum_o Y_p[a,o] PO[b,o]          for b in the fragment of row a
//   J[(f;a,b),p] = 2 (G_p[a,b] + G_p[b,a])
#include "common.cuh"

namespace qcdmet {

namespace {
constexpr int TA = 64;   // fragment rows per CTA
constexpr int TO = 64;   // occupied orbitals per tile
constexpr int TV = 32;   // virtual orbitals per chunk
constexpr int NT = 256;  // threads per CTA: 16 x 16, each a 4 x 4 register tile of Y
constexpr int PVS_LD = TA + 2, ZS_LD = TO + 2, YS_LD = TO + 1;
constexpr int GEN_SMEM = (TV * PVS_LD + TV * ZS_LD + TA * YS_LD) * 8 + (TA + 1) * 4;
}

__global__ void __launch_bounds__(NT)
jac_general_kernel(JacView v, double *__restrict__ G)
{
    extern __shared__ __align__(16) unsigned char gsm[];
    double *PVs = reinterpret_cast<double *>(gsm);          // PVs[vv][a]
    double *Zs = PVs + TV * PVS_LD;                          // Zs[vv][o]
    double *Ys = Zs + TV * ZS_LD;                            // Ys[a][o]
    int *pairOff = reinterpret_cast<int *>(Ys + TA * YS_LD);

    const int tid = threadIdx.x;
    const int a0 = blockIdx.x * TA;
    const int pl = blockIdx.y;                 // local column
    const int p = v.colBegin + pl;
    const int eBeg = v.h1start[p], eEnd = v.h1start[p + 1];
    double *Gp = G + (size_t)v.Q * pl;

    if (tid == 0) {
        int acc = 0;
        for (int a = 0; a < TA; ++a) {
            pairOff[a] = acc;
            int ga = a0 + a;
            if (ga < v.S) { int f = v.rowFrag[ga]; acc += v.rowStart[f + 1] - v.rowStart[f]; }
        }
        pairOff[TA] = acc;
    }
    __syncthreads();
    const int npairs = pairOff[TA];

    const int ty = tid / 16, tx = tid % 16;    // Y register tile: rows 4 ty .. 4 ty + 3, cols 4 tx .. 4 tx + 3
    const int zo = tid % TO, zv0 = tid / TO;   // Z build: column zo, rows zv0 + 4j

    for (int o0 = 0; o0 < v.no; o0 += TO) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

        const int og = o0 + zo;

        for (int v0 = 0; v0 < v.nv; v0 += TV) {
            // stage PV[a-tile][v-chunk]
            for (int t = tid; t < TA * TV; t += NT) {
                int a = t / TV, vv = t % TV;
                int ga = a0 + a, gv = v0 + vv;
                PVs[vv * PVS_LD + a] = (ga < v.S && gv < v.nv) ? v.PV[(size_t)ga * v.ldv + gv] : 0.0;
            }
            // build the Z tile from the CSR elements of the perturbation
#pragma unroll
            for (int j = 0; j < TV / (NT / TO); ++j) {
                int vv = zv0 + (NT / TO) * j;
                int gv = v0 + vv;
                double z = 0.0;
                if (gv < v.nv && og < v.no) {
                    double s = 0.0;
                    for (int el = eBeg; el < eEnd; ++el)
                        s += v.CvT[(size_t)v.h1row[el] * v.ldv + gv] *
                             v.CoT[(size_t)v.h1col[el] * v.ldo + og];
                    z = s * v.W[(size_t)gv * v.ldo + og];
                }
                Zs[vv * ZS_LD + zo] = z;
            }
            __syncthreads();
#pragma unroll 8
            for (int vv = 0; vv < TV; ++vv) {
                const double2 pa = *reinterpret_cast<const double2 *>(PVs + vv * PVS_LD + 4 * ty);
                const double2 pb = *reinterpret_cast<const double2 *>(PVs + vv * PVS_LD + 4 * ty + 2);
                const double2 za = *reinterpret_cast<const double2 *>(Zs + vv * ZS_LD + 4 * tx);
                const double2 zb = *reinterpret_cast<const double2 *>(Zs + vv * ZS_LD + 4 * tx + 2);
                const double pr[4] = {pa.x, pa.y, pb.x, pb.y}, zc[4] = {za.x, za.y, zb.x, zb.y};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += pr[i] * zc[j];
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) Ys[(4 * ty + i) * YS_LD + 4 * tx + j] = acc[i][j];
        __syncthreads();
        // epilogue: contract with the occupied panel rows of the same fragment
        const int ow = min(TO, v.no - o0);
        for (int t = tid; t < npairs; t += NT) {
            int lo = 0, hi = TA;
            while (hi - lo > 1) { int mid = (lo + hi) / 2; if (pairOff[mid] <= t) lo = mid; else hi = mid; }
            const int a = lo, ga = a0 + a;
            const int f = v.rowFrag[ga], rs = v.rowStart[f], rf = v.rowStart[f + 1] - rs;
            const int bl = t - pairOff[a];
            const double *pb = v.PO + (size_t)(rs + bl) * v.ldo + o0;
            double s = 0.0;
            for (int oo = 0; oo < ow; ++oo) s += Ys[a * YS_LD + oo] * pb[oo];
            double *dst = Gp + v.qoff[f] + (ga - rs) + (long long)rf * bl;
            if (o0 == 0) *dst = s; else *dst += s;
        }
        __syncthreads();
    }
}

// In place: J[(f;a,b)] = 2 (G[(f;a,b)] + G[(f;b,a)])   (lib/rhf_response.cc:104-108 on blocks)
__global__ void symmetrise_kernel(JacView v, double *__restrict__ G)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= v.Q) return;
    double *Gp = G + (size_t)v.Q * blockIdx.y;
    int lo = 0, hi = v.F;
    while (hi - lo > 1) { int mid = (lo + hi) / 2; if (v.qoff[mid] <= q) lo = mid; else hi = mid; }
    const int f = lo, rf = v.rowStart[f + 1] - v.rowStart[f];
    const long long loc = q - v.qoff[f];
    const int a = (int)(loc % rf), b = (int)(loc / rf);
    if (a > b) return;
    double *x = Gp + v.qoff[f] + a + (long long)rf * b;
    double *y = Gp + v.qoff[f] + b + (long long)rf * a;
    const double s = 2.0 * (*x + *y);
    *x = s;
    *y = s;
}

int launch_jacobian_general(const JacView &v, cudaStream_t st, int *launches, double *flops,
                            cudaEvent_t ev0, cudaEvent_t ev1)
{
    const int ncols = v.colEnd - v.colBegin;
    *launches = 0;
    *flops = 0.0;
    if (ncols <= 0 || v.S == 0) return QCDMET_OK;
    const int atiles = (v.S + TA - 1) / TA;
    // per device (function attributes live in the context of the current device): set on every launch
    QC_CUDA(cudaFuncSetAttribute(jac_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEN_SMEM));
    // gridDim.y is limited to 65535: chunk the columns
    for (int c0 = 0; c0 < ncols; c0 += 32768) {
        const int nc = min(32768, ncols - c0);
        JacView w = v;
        w.colBegin = v.colBegin + c0;
        w.colEnd = w.colBegin + nc;
        double *G = v.J + (size_t)v.Q * c0;
        if (c0 == 0 && ev0) QC_CUDA(cudaEventRecord(ev0, st));
        jac_general_kernel<<<dim3(atiles, nc), NT, GEN_SMEM, st>>>(w, G);
        QC_CUDA(cudaGetLastError());
        if (c0 == 0 && ev1) QC_CUDA(cudaEventRecord(ev1, st));
        const long long qb = (v.Q + 255) / 256;
        symmetrise_kernel<<<dim3((unsigned)qb, nc), 256, 0, st>>>(w, G);
        QC_CUDA(cudaGetLastError());
        *launches += 2;
    }
    *flops = -1.0;   // filled by the caller from the CSR (host-side count)
    return QCDMET_OK;
}

}  // namespace qcdmet
