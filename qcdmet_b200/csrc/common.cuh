// common.cuh -- shared declarations of libqcdmet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/qcdmet_b200.h"

namespace qcdmet {

void set_error(const char *fmt, ...);

#define QC_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            qcdmet::set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__,           \
                              (int)err__, cudaGetErrorString(err__), #call);                  \
            return QCDMET_E_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define QC_CHECK(cond, code, ...)                                                             \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            qcdmet::set_error(__VA_ARGS__);                                                   \
            return (code);                                                                    \
        }                                                                                     \
    } while (0)

#define QC_TRY(expr)                                                                          \
    do {                                                                                      \
        int rc__ = (expr);                                                                    \
        if (rc__ != QCDMET_OK) return rc__;                                                   \
    } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Debug build (make debug -> libqcdmet_b200_dbg.so, -DQCDMET_BOUNDS_CHECK): every Jacobian / panel store and the
// table-driven shared-memory indices of the kernels are range-checked on the device; a violation prints where and
// traps (the next CUDA call returns an error).  compute-sanitizer is not available on the GPU pool.
#ifdef QCDMET_BOUNDS_CHECK
#define QC_DEV_ASSERT(cond)                                                                              \
    do {                                                                                                 \
        if (!(cond)) {                                                                                   \
            printf("qcdmet bounds check failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, \
                   (int)blockIdx.x, (int)threadIdx.x);                                                   \
            __trap();                                                                                    \
        }                                                                                                \
    } while (0)
#else
#define QC_DEV_ASSERT(cond) ((void)0)
#endif

// ---------------------------------------------------------------------------------------
// Device-side view of everything the Jacobian kernels read.  Layout in HBM:
//   CoT [N ][ldo]  CoT[i][o] = C[i, o]          (site-major, orbital contiguous, zero padded)
//   CvT [N ][ldv]  CvT[i][v] = C[i, no + v]
//   PO  [S ][ldo]  fragment rows projected on the occupied orbitals (R*C_o or gathered CoT)
//   PV  [S ][ldv]  fragment rows projected on the virtual orbitals
//   e   [N]        eigenvalues, ascending
//   J   [ncols][Q] column-major (Q, ncols): one Jacobian column per perturbation
// ---------------------------------------------------------------------------------------
struct JacView {
    int N, no, nv, ldo, ldv;
    int P;                 // perturbations in the plan
    int colBegin, colEnd;  // column shard
    int F;                 // fragments
    int S;                 // fragment rows
    long long Q;           // residuals
    const int *h1start, *h1row, *h1col;
    const int *rowStart;       // [F+1]
    const int *rowFrag;        // [S] fragment of each row
    const long long *qoff;     // [F+1] residual offset of each fragment
    const double *CoT, *CvT, *PO, *PV, *e;
    const double *W;           // [nv][ldo] energy denominators 1/(e_o - e_v) (general kernel)
    double *J;                 // [Q * ncols]
};

// kernels_prep.cu
int launch_split_transpose(const double *C, int N, int no, double *CoT, int ldo, double *CvT,
                           int ldv, cudaStream_t st);
int launch_gather_rows(const double *src, int ld, const int *rowIdx, int S, double *dst,
                       cudaStream_t st);
int launch_dblocks(const JacView &v, double *D, cudaStream_t st);
int launch_one_rdm(const double *CoT, int ldo, int N, int no, double *D, cudaStream_t st);
int launch_build_w(const double *e, int no, int nv, int ldo, double *W, cudaStream_t st);

// kernels_general.cu
// ev0/ev1 are recorded on `st` around the dominant kernel (first column chunk).
int launch_jacobian_general(const JacView &v, cudaStream_t st, int *launches, double *flops,
                            cudaEvent_t ev0, cudaEvent_t ev1);

// kernels_bath.cu: Schmidt baths of all fragments from the occupied orbitals (m x m Jacobi per fragment + one DGEMM)
#define QCDMET_BATH_MAX_IMP 64          // impurity orbitals per fragment the batched bath kernel handles
int launch_construct_baths(cublasHandle_t blas, const double *CoT, int ldo, int N, int no, int F, int S, int maxM,
                           const int *d_impStart, const int *d_impIdx, const int *d_rowFrag, const int *d_numReq,
                           double threshold, double *d_Y, double *d_occ, int *d_numKept, int *d_sweeps, double *d_B,
                           cudaStream_t st);

#define QCDMET_EMBED_MAX_TEI_ORBS 16     // embedded orbitals per fragment up to which the two-electron block is returned
int launch_embed_hubbard(cublasHandle_t blas, const double *CoT, int ldo, int N, int no, int F, int S, const double *d_H, double U,
                         const int *d_impStart, const int *d_impIdx, const int *d_numKept, const int *d_cap,
                         const long long *d_off2, const long long *d_off4, int nbMax, const double *d_B, const double *d_occ, double *d_HB,
                         double *d_ddiag, double *d_oei, double *d_fock, double *d_tei, double *d_coreTrace, cudaStream_t st);

// Destinations of the fused all-gather: every rank of a node maps every rank's full Jacobian buffer;
// the DMMA kernel stores each element it produces into all of them (NVLink peer stores).
#define QCDMET_MAX_PEERS 8
struct PeerSet {
    int nranks = 1, rank = 0;
    double *J[QCDMET_MAX_PEERS] = {nullptr};
};

// kernels_dmma.cu
struct DmmaPlanData;   // site-factorised schedule (device arrays), built once per plan
bool dmma_eligible(int N, int P, int F, const int *h1start, const int *h1row, const int *h1col,
                   const int *rowStart, const int *rowIdx, bool denseRows, std::string *why);
int dmma_plan_build(DmmaPlanData **out, int N, int no, int P, int F, const int *h1start,
                    const int *h1row, const int *h1col, const int *rowStart, const int *rowIdx);
void dmma_plan_free(DmmaPlanData *d);
int dmma_prepare_orbitals(DmmaPlanData *d, const JacView &v, cudaStream_t st);
int launch_jacobian_dmma(DmmaPlanData *d, const JacView &v, cudaStream_t st, int *launches,
                         double *flops, cudaEvent_t ev0, cudaEvent_t ev1, bool allowSym,
                         const PeerSet *peers);
bool dmma_last_launch_symmetric(const DmmaPlanData *d);
bool dmma_last_launch_split(const DmmaPlanData *d);      // tile-granular balancing (SPLIT kernel) was used
bool dmma_plan_expanded(const DmmaPlanData *d);   // the list was split into per-cell sub-perturbations (TI)
void dmma_schedule_info(int N, int no, int P, int F, const int *h1start, const int *h1row, const int *h1col,
                        const int *rowStart, const int *rowIdx, int *info, std::string *why);

// kernels_contract.cu: per-perturbation contraction of the site panels (factorised path)
#define QCDMET_CONTRACT_MAX_NNZ 64      // elements per perturbation the contraction kernel stages
#define QCDMET_CONTRACT_STRIPS_PER_WARP 3
struct ContractSchedule {               // host-built, see kernels_contract.cu
    int nTiles = 0;
    long long validSlots = 0;           // row slots that hold a fragment row (flop count)
    long long blockPairs = 0;           // 8 x 8 block pairs over all tiles (flop count)
    std::vector<int> tileInfo;          // [nTiles][4] slotBegin, nSlots, first entry in strips, nStrips
    std::vector<int> slotRow;           // fragment row of every slot, -1 = padding of a ragged block
    std::vector<int> strips;            // [nTiles][8 warps][3] b1 | b2 << 8 | n2 << 16 (tile-local blocks), -1 = none
};
struct ContractPlan {
    ContractSchedule sched;             // tiles of <= 16 blocks (chunked kernel)
    int *d_tileInfo = nullptr, *d_slotRow = nullptr, *d_strips = nullptr;
    ContractSchedule sched8;            // tiles of <= 8 blocks (resident kernel)
    int *d_tileInfo8 = nullptr, *d_slotRow8 = nullptr, *d_strips8 = nullptr;
    // resident variant (small occupied spaces): CTA table (tile, first column, columns) for the column count resCols
    void *d_ctaTab = nullptr;
    size_t ctaCapacity = 0;
    int nCtas = 0, resCols = -1;
    bool res64 = false;                 // the table was built for the 64-slot variant
    int maxFragBlocks = 0;              // 8-row blocks of the largest fragment
};
void contract_schedule_build(int F, const int *rowStart, ContractSchedule *out, int maxBlocks = 0);   // 0: 16 blocks per tile
int contract_plan_build(ContractPlan **out, int F, const int *rowStart);
void contract_plan_free(ContractPlan *c);
int launch_contract_panels(ContractPlan *c, const JacView &v, const double *M, const int *elU,
                           long long nnzInRange, bool nnzAtMost2, cudaStream_t st, int *launches, double *flops);

// factorised path for large fragments (kernels_dmma.cu panel mode + kernels_contract.cu)
struct FactPlanData;
bool fact_eligible(int P, const int *h1start, std::string *why);
size_t fact_plan_bytes(int N, int no, int P, int S, const int *h1start, const int *h1row);
int fact_plan_build(FactPlanData **out, int N, int no, int P, int F, const int *rowStart, int ldo,
                    const int *h1start, const int *h1row);
void fact_plan_free(FactPlanData *f);
int fact_prepare_orbitals(FactPlanData *f, const JacView &v, cudaStream_t st);
int launch_jacobian_factorised(FactPlanData *f, const JacView &v, long long nnzInRange, cudaStream_t st,
                               int *launches, double *flops, cudaEvent_t ev0, cudaEvent_t ev1);

}  // namespace qcdmet
