/* qcdmet_b200.h -- C ABI of libqcdmet_b200.so (sm_100a).
 *
 * Drop-in boundary for QC-DMET's RHF 1-RDM linear response hot path.  The reference binds
 * its native library with ctypes (src/qcdmethelper.py:23-24) and calls exactly one symbol,
 *     extern "C" void rhf_response(...)                       lib/rhf_response.cc:31-32
 * from qcdmethelper.construct1RDM_response (src/qcdmethelper.py:102-109); the Jacobian of
 * the umat fit is then assembled in Python from the N x N x P response
 * (dmet.rdm_differences_derivative, src/dmet.py:477-525).
 *
 * This library exports
 *   (1) rhf_response           -- same symbol, same arguments, same semantics (compat);
 *   (2) qcdmet_jacobian*       -- the fused path: fragment blocks of dD_p written directly
 *                                 as Jacobian columns, replacing lib/rhf_response.cc:73-112
 *                                 + src/dmet.py:496-520 in one call;
 *   (3) qcdmet_plan_*          -- the same path with device-resident state (used by the
 *                                 Python mirror of qcdmethelper/dmet and by bench.py).
 *
 * Plain pointers and sizes only.  All entry points return 0 on success, a negative
 * QCDMET_E_* code otherwise; qcdmet_last_error() gives the message.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with QCDMET_E_CUDA.
 *
 * Array contracts (identical to the reference's, src/qcdmethelper.py:62-64,97,100):
 *   H1start[Nterms+1] monotone, H1start[0] = 0; H1row/H1col 0-based site indices < Norb,
 *   int32; every stored element has value 1.  H0: symmetric Norb x Norb float64.
 */
#ifndef QCDMET_B200_H
#define QCDMET_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCDMET_OK          0
#define QCDMET_E_ARG      -1   /* invalid argument                                   */
#define QCDMET_E_CUDA     -2   /* CUDA runtime / no device                           */
#define QCDMET_E_SOLVER   -3   /* cuSOLVER syevd failed (info != 0)                  */
#define QCDMET_E_ALLOC    -4   /* device or host allocation failed                   */
#define QCDMET_E_STATE    -5   /* plan used before orbitals / Jacobian are available */
#define QCDMET_E_GAP      -6   /* e[numPairs] - e[numPairs-1] <= 0: response diverges */

/* Kernel selection for qcdmet_plan_set_kernel(). */
#define QCDMET_KERNEL_AUTO     0   /* fused DMMA kernel when eligible, else factorised, else general */
#define QCDMET_KERNEL_GENERAL  1   /* plain FP64 general kernel (any CSR, any rows)           */
#define QCDMET_KERNEL_DMMA     2   /* force the DMMA kernel (error if not eligible)           */
#define QCDMET_KERNEL_DMMA_NOSYM 3 /* DMMA kernel without the symmetric (mirrored-patch) mode */
#define QCDMET_KERNEL_FACTORISED 4 /* site panels M_r by the DMMA main loop into HBM + per-perturbation
                                      contraction: any fragment sizes, <= 64 elements per perturbation,
                                      panels must fit in device memory (error otherwise)                */

/* ---------------------------------------------------------------------------------------
 * (1) Compat symbol.  Replaces lib/rhf_response.cc:32-119 one to one:
 *     diagonalises H0 (cuSOLVER syevd instead of dsyev, :42-52), overwrites H0 with the
 *     1-RDM 2*C_o*C_o^T (:57-64), and fills rdm_deriv[row + Norb*(col + Norb*deriv)] with
 *     dD_deriv (:104-108).  The reference returns void; an int status is ABI-compatible with
 *     its ctypes call site (src/qcdmethelper.py:102), which ignores the return value.
 *     Unlike the reference, indices are 64-bit: Norb^2*Nterms >= 2^31 is legal.
 * ------------------------------------------------------------------------------------- */
int rhf_response(const int Norb, const int Nterms, const int numPairs,
                 int *H1start, int *H1row, int *H1col, double *H0, double *rdm_deriv);

/* ---------------------------------------------------------------------------------------
 * (2) Fused Jacobian, host buffers in / host buffers out.
 *     Replaces the pair  helper.construct1RDM_response (src/qcdmethelper.py:84-112)  +
 *     the assembly loop of dmet.rdm_differences_derivative (src/dmet.py:484-520).
 *
 *     Fragments: Nfrag blocks of residual rows, fragment f owning rows
 *     rowStart[f] .. rowStart[f+1]-1 (S = rowStart[Nfrag] rows in total).
 *       R != NULL : dense projector rows, row-major S x Norb; the rows of fragment f are
 *                   dmetOrbs[f]^T (fitImpBath == True, src/dmet.py:502)
 *       R == NULL : rowIdx[S] site index per row = flatnonzero(impClust[f]) concatenated
 *                   (fitImpBath == False, src/dmet.py:509)
 *     Output: J[q + Q*(p - colBegin)], Q = sum_f r_f^2, p in [colBegin, colEnd); inside
 *     fragment f the r_f x r_f block of dD_p is stored column-major (reshape order='F',
 *     src/dmet.py:515-517).  This is exactly the memory of the (Q, P) array the reference
 *     returns (np.array(list).T, src/dmet.py:520).
 *     D_blocks (optional, may be NULL): Q doubles, the same blocks of the mean-field 1-RDM
 *     2*C_o*C_o^T (the model part of dmet.rdm_differences, src/dmet.py:416-421).
 *     eig_out (optional): Norb eigenvalues, ascending.
 * ------------------------------------------------------------------------------------- */
int qcdmet_jacobian(int Norb, int Nterms, int numPairs,
                    const int *H1start, const int *H1row, const int *H1col,
                    const double *H0,
                    int Nfrag, const int *rowStart, const double *R, const int *rowIdx,
                    int colBegin, int colEnd,
                    double *J, double *D_blocks, double *eig_out);

/* qcdmet_jacobian* keep the plan (device buffers, schedule, library handles) of their last call and reuse it
 * when the perturbation list and fragment rows are unchanged; qcdmet_shutdown() releases it. */
int qcdmet_shutdown(void);

/* Test-only variant: orbitals given instead of H0 (no eigensolve), so that the kernels
 * can be compared against the oracle on bit-identical (C, e).
 * C: column-major Norb x Norb, eigenvector k in column k (LAPACK layout), e ascending. */
int qcdmet_jacobian_from_orbitals(int Norb, int Nterms, int numPairs,
                    const int *H1start, const int *H1row, const int *H1col,
                    const double *C, const double *e,
                    int Nfrag, const int *rowStart, const double *R, const int *rowIdx,
                    int colBegin, int colEnd,
                    double *J, double *D_blocks);

/* ---------------------------------------------------------------------------------------
 * (3) Plans: perturbation list + fragment rows live on the device across calls
 *     (the reference freezes list_H1 at dmet.__init__, src/dmet.py:86).
 * ------------------------------------------------------------------------------------- */
typedef struct qcdmet_plan qcdmet_plan;

/* device < 0: use the current CUDA device.  Copies all index arrays and R to the device. */
int qcdmet_plan_create(qcdmet_plan **plan, int device, int Norb, int Nterms, int numPairs,
                       const int *H1start, const int *H1row, const int *H1col,
                       int Nfrag, const int *rowStart, const double *R, const int *rowIdx);
int qcdmet_plan_destroy(qcdmet_plan *plan);

/* ---- Single-process multi-GPU: the in-library parallel region of lib/rhf_response.cc:73-79 -------------
 * The reference parallelises inside the library call (an OpenMP team over the perturbations); here a plan
 * can span several GPUs of one node the same way, so the ctypes boundary stays a drop-in:
 *   qcdmet_set_devices(n)        process-wide default for plans created afterwards on the current device
 *                                (device < 0 in qcdmet_plan_create, i.e. also qcdmet_jacobian* and the Python
 *                                helper); n = -1: every visible device.  $QCDMET_B200_DEVICES = n | "all" sets
 *                                the same default without touching the caller's code.
 *   qcdmet_plan_set_devices      the same for one plan (devices plan->device, +1, ... mod the device count).
 * One eigensolve on the first device, (C, e) copied to the others over NVLink (bit-identical orbitals);
 * every device launches its share from the caller's thread; an event barrier replaces the collective.
 * Fused DMMA kernel: each device executes 1/n of the block-pair list and stores every Jacobian element into all
 * devices' buffers (NVLink peer stores).  Other kernels: contiguous column slices, copied to the first device.
 * qcdmet_plan_get_jacobian ships 1/n of the columns from every device over its own PCIe link. */
int qcdmet_set_devices(int n);
int qcdmet_get_devices(void);
int qcdmet_plan_set_devices(qcdmet_plan *plan, int n);
int qcdmet_plan_num_devices(qcdmet_plan *plan);

/* Page-locked host memory (cudaHostAlloc, portable): destination buffers that the device->host copy of the
 * Jacobian reaches at full PCIe rate from every device. */
void *qcdmet_host_alloc(size_t bytes);
int qcdmet_host_free(void *ptr);

/* Restrict the plan to Jacobian columns [colBegin, colEnd) (multi-GPU column sharding). */
int qcdmet_plan_set_columns(qcdmet_plan *plan, int colBegin, int colEnd);
/* QCDMET_KERNEL_*; default AUTO. */
int qcdmet_plan_set_kernel(qcdmet_plan *plan, int kernel);
/* CUDA stream (cudaStream_t) every launch of this plan goes to; NULL = legacy default stream. */
int qcdmet_plan_set_stream(qcdmet_plan *plan, void *stream);

/* Write the Jacobian slice into caller-owned device memory (e.g. this rank's slot of an NCCL
 * all-gather buffer) instead of the plan's own buffer; capacity in doubles.  NULL restores the
 * plan-owned buffer. */
int qcdmet_plan_set_output_device(qcdmet_plan *plan, void *device_ptr, long long capacity);

/* ---- Multi-GPU, one process per GPU on one node: fused compute + all-gather over NVLink --------
 * Every rank exports the IPC handle of its full (Q x P) Jacobian buffer, the handles are exchanged
 * by any host transport (bench.py: torch.distributed.all_gather_object), and every rank maps all
 * peers.  compute() then executes this rank's share of the (equal-cost) block-pair work list and
 * stores every Jacobian element it produces into ALL ranks' buffers with peer stores; after a
 * cross-rank barrier placed behind compute() on the stream, every rank holds the complete Jacobian.
 * Needs the DMMA kernel and the full column range. */
#define QCDMET_IPC_HANDLE_BYTES 64
#define QCDMET_MAX_RANKS 8
int qcdmet_plan_export_jacobian(qcdmet_plan *plan, void *handle /* QCDMET_IPC_HANDLE_BYTES */);
int qcdmet_plan_attach_peers(qcdmet_plan *plan, int nranks, int rank, const void *handles /* nranks x 64 bytes */);
/* Same with raw device pointers (single-process multi-plan set-ups and tests); ptrs[rank] may be NULL. */
int qcdmet_plan_attach_peer_pointers(qcdmet_plan *plan, int nranks, int rank, void *const *ptrs);
int qcdmet_plan_detach_peers(qcdmet_plan *plan);

/* Upload H0 (host), run cuSOLVER syevd on the device, build the operand panels. */
int qcdmet_plan_set_h0(qcdmet_plan *plan, const double *H0);
/* Upload orbitals (host; layout as in qcdmet_jacobian_from_orbitals), build the panels. */
int qcdmet_plan_set_orbitals(qcdmet_plan *plan, const double *C, const double *e);

/* Device pointers of the plan's orbital buffers (C: column-major Norb x Norb, eigenvector k in column k; e:
 * ascending) so that a caller can fill them in place -- e.g. an NCCL broadcast from the rank that ran the
 * eigensolve (set_h0) -- and then, after synchronising its own stream, commit them (panels are rebuilt). */
int qcdmet_plan_orbitals_device(qcdmet_plan *plan, void **C, void **e);
int qcdmet_plan_commit_orbitals(qcdmet_plan *plan);

/* Launch the Jacobian kernels for the plan's column range (asynchronous on the plan stream;
 * inputs and output stay in HBM). */
int qcdmet_plan_compute(qcdmet_plan *plan);
/* Synchronise the plan stream; returns the first asynchronous error, if any. */
int qcdmet_plan_sync(qcdmet_plan *plan);

/* Device pointer of the column slice J[Q x (colEnd-colBegin)], column-major (for NCCL). */
void *qcdmet_plan_jacobian_device(qcdmet_plan *plan);
/* Blocking device->host copies. */
int qcdmet_plan_get_jacobian(qcdmet_plan *plan, double *J);
/* Columns [c0, c1) of the computed slice only (J points at the first of them). */
int qcdmet_plan_get_jacobian_columns(qcdmet_plan *plan, int c0, int c1, double *J);
int qcdmet_plan_get_dblocks(qcdmet_plan *plan, double *D_blocks);
int qcdmet_plan_get_eigenvalues(qcdmet_plan *plan, double *e);
int qcdmet_plan_get_orbitals(qcdmet_plan *plan, double *C);
/* Full response written like the compat symbol (small problems): rdm_deriv[N*N*ncols], and
 * the 1-RDM (N*N) if one_rdm != NULL. */
int qcdmet_plan_get_full_response(qcdmet_plan *plan, double *rdm_deriv, double *one_rdm);

/* Mean-field 1-RDM 2 C_o C_o^T of the plan's orbitals, Norb x Norb doubles (symmetric); the matrix
 * qcdmethelper.construct1RDM_loc returns (src/qcdmethelper.py:67-82,114-122). */
int qcdmet_plan_get_one_rdm(qcdmet_plan *plan, double *one_rdm);

/* Schmidt baths of Nfrag impurities in one call, from the plan's current orbitals.  Replaces
 * qcdmethelper.constructbath (src/qcdmethelper.py:124-168: numpy eigh of the (N-m) x (N-m) environment block per
 * fragment) by one m x m eigenproblem per fragment + one DGEMM, exact for the idempotent mean-field 1-RDM.
 *   impStart[Nfrag+1], impIdx[S]: impurity orbitals of fragment f are impIdx[impStart[f] .. impStart[f+1]), m_f <= 64
 *   numBathReq[Nfrag] (or NULL = m_f): bath orbitals wanted; threshold: the reference's 1e-13 (:124)
 *   numBath[Nfrag]  out: bath orbitals kept = min(numBathReq, #candidates with min(occ, 2 - occ) > threshold) (:141-144)
 *   occ[S]          out (may be NULL): environment occupations of the m_f candidates of each fragment, in rank order
 *                   (closest to 1 first, :139-140); the first numBath[f] belong to the bath orbitals
 *   bath[S * Norb]  out: row impStart[f] + k = k-th bath orbital of fragment f over the Norb lattice orbitals (zero on
 *                   the fragment's own impurity orbitals, unit norm); rows k >= numBath[f] are zero.
 * The reference's loc2dmet[:, :m+nb] is [unit vectors of the impurity orbitals | these rows^T]; eigenvector signs are
 * arbitrary in both. */
int qcdmet_plan_construct_baths(qcdmet_plan *plan, int Nfrag, const int *impStart, const int *impIdx,
                                const int *numBathReq, double threshold, int *numBath, double *occ, double *bath);

/* Baths AND embedded Hamiltonians of all fragments for an on-site (Hubbard) interaction U sum_x n_x,up n_x,down, in one
 * call: what the reference's DMET iteration does per fragment with constructbath + localintegrals_hubbard.dmet_oei /
 * dmet_fock / dmet_tei (src/dmet.py:194-214, src/localintegrals_hubbard.py:88-116; dmet_tei is an O(n^4 N) Python loop).
 *   hopping[Norb*Norb]: the one-electron matrix (symmetric); impStart/impIdx/numBathReq/threshold/numBath/occ/bath as in
 *   qcdmet_plan_construct_baths (bath may be NULL).
 *   Fragment f has cap_f = m_f + min(numBathReq_f, m_f) embedded-orbital slots (impurity orbitals first, then bath orbitals);
 *   slots of bath orbitals that were not kept are zero.  Blocks are packed fragment after fragment, row-major:
 *   oei, fock: cap_f^2 doubles each (c^T h c; c^T (h + U/2 diag(core 1-RDM)) c = loc_rhf_fock_bis of the core density);
 *   tei (or NULL; cap_f <= 16): cap_f^4 doubles, tei[i,j,k,l] = U sum_x c_xi c_xj c_xk c_xl;
 *   coreElectrons (or NULL): trace of the core 1-RDM per fragment (Nelec_in_imp = Nelec - round(that), src/dmet.py:207). */
int qcdmet_plan_embed_hubbard(qcdmet_plan *plan, const double *hopping, double U, int Nfrag, const int *impStart,
                              const int *impIdx, const int *numBathReq, double threshold, int *numBath, double *occ,
                              double *bath, double *oei, double *fock, double *tei, double *coreElectrons);

/* Levenberg-Marquardt step from the Jacobian resident on the device (no 8 Q P byte D2H):
 *   (J^T J + lambda diag(J^T J)) step = -J^T resid        resid: Q doubles (host), step: ncols doubles (host)
 * cuBLAS DSYRK/DGEMV + cuSOLVER potrf/potrs; J^T J is kept until the next compute, so retries with a
 * different lambda cost one Cholesky.  info (optional, 2 doubles): max|J^T resid|, (J^T resid).step.
 * Replaces the MINPACK lmder iteration behind scipy.optimize.leastsq (src/dmet.py:631) for large fits.
 * QCDMET_E_SOLVER if the damped matrix is not positive definite (raise lambda). */
int qcdmet_plan_lm_step(qcdmet_plan *plan, const double *resid, double lambda, double *step, double *info);
/* How qcdmet_plan_lm_step solves the damped problem.  QCDMET_LM_NORMAL (default): equilibrated normal equations,
 * Cholesky; J^T J is formed once per Jacobian (sharded over the devices of a multi-device plan) and a retry with
 * another lambda costs one Cholesky.  QCDMET_LM_QR: Householder QR (cuSOLVER geqrf/ormqr) of the augmented matrix
 * [J; sqrt(lambda) D] -- what MINPACK's lmpar factorises -- for fits whose cond(J)^2 exceeds double precision;
 * needs (Q + P) P doubles of scratch and one QR per lambda. */
#define QCDMET_LM_NORMAL 0
#define QCDMET_LM_QR 1
int qcdmet_plan_set_lm_method(qcdmet_plan *plan, int method);

/* grad = 2 J^T resid from the device-resident Jacobian (the BFGS gradient of src/dmet.py:381-388);
 * resid: Q doubles (host), grad: ncols doubles (host). */
int qcdmet_plan_gradient(qcdmet_plan *plan, const double *resid, double *grad);

/* Introspection (bench / tests). */
long long qcdmet_plan_num_residuals(qcdmet_plan *plan);    /* Q                                  */
int  qcdmet_plan_kernel_used(qcdmet_plan *plan);           /* QCDMET_KERNEL_GENERAL, _DMMA or _FACTORISED */
int  qcdmet_plan_launches(qcdmet_plan *plan);              /* kernels launched by the last compute */
int  qcdmet_plan_symmetric_used(qcdmet_plan *plan);        /* 1 if the last compute used the mirrored-patch mode */
/* 1 if the last fused launch balanced the SMs at column-tile granularity (an item cut by a range boundary is finished by
 * the CTA holding its first tiles, in tile order: results are bit-identical to the unsplit launch); $QCDMET_B200_SPLIT=0|1
 * forces it off / on (tuning, tests). */
int  qcdmet_plan_split_used(qcdmet_plan *plan);
double qcdmet_plan_flops(qcdmet_plan *plan);               /* FP64 flops executed by the last compute */
/* Device time of the dominant kernel of the last compute (CUDA events recorded on the plan
 * stream around that launch; synchronises the stream).  ms < 0 if nothing was computed. */
int  qcdmet_plan_kernel_time_ms(qcdmet_plan *plan, double *ms);

/* Breakdown of the last step in ms (-1 = not measured): ms[5] = { H2D of H0, cuSOLVER syevd, panel build
 * (CUDA events inside qcdmet_plan_set_h0), dominant kernel of the first device's share, device->host copy of the
 * Jacobian (wall clock of the last qcdmet_plan_get_jacobian*) }.  Synchronises the plan. */
int  qcdmet_plan_timings(qcdmet_plan *plan, double *ms);
/* Checksum of the computed slice on the device: out[2] = { sum, sum of squares } in a fixed reduction order
 * (reproducible run to run and device to device), *bits = xor of the 64-bit patterns of all elements.
 * identical (may be NULL): 1 if every device of a multi-device plan holds a bit-identical Jacobian. */
int  qcdmet_plan_jacobian_checksum(qcdmet_plan *plan, double *out, unsigned long long *bits, int *identical);
/* Benchmark utility: overwrite a 256 MB scratch buffer on every device of the plan (on the plan's streams), so
 * that the next step does not start from a warm 126 MB L2. */
int  qcdmet_plan_flush_l2(qcdmet_plan *plan);
/* Note left by QCDMET_KERNEL_AUTO when it had to use the general FP64 kernel (why the fused and the
 * factorised path do not apply); empty string otherwise. */
const char *qcdmet_plan_info(qcdmet_plan *plan);

/* Sustained DMMA.8x8x4 (mma.sync.m8n8k4.f64) issue rate of a device in TFLOP/s, measured for about `seconds`:
 * the FP64 tensor-pipe ceiling the roofline of the Jacobian kernel is quoted against (MEASURED_PEAKS.json has no
 * FP64 entry).  device < 0: current device. */
int qcdmet_measure_fp64_peak(int device, double seconds, double *tflops);

/* Host-only (no device needed): the schedule the DMMA kernel would use for these lists.
 * info[10] = { eligible, rows-per-fragment variant (4|8), 8-column fragments per warp (8|6|4), components,
 *              row tiles, symmetric mode possible, list was split (TI), internal columns, block pairs of a
 *              full-range launch, column tiles }.  rowIdx == NULL means dense projector rows. */
int qcdmet_schedule_info(int Norb, int Nterms, int numPairs, const int *H1start, const int *H1row,
                         const int *H1col, int Nfrag, const int *rowStart, const int *rowIdx, int *info);

/* Host-only: the kernel QCDMET_KERNEL_AUTO selects for these lists (QCDMET_KERNEL_DMMA, _FACTORISED or _GENERAL)
 * and, for the factorised path, the device memory its site panels and tiled operands need (panel_bytes, may be
 * NULL).  At plan creation the factorised path is additionally refused -- and the general kernel used -- when
 * panel_bytes exceeds half of the free device memory.  rowIdx == NULL means dense projector rows. */
int qcdmet_auto_kernel(int Norb, int Nterms, int numPairs, const int *H1start, const int *H1row, const int *H1col,
                       int Nfrag, const int *rowStart, const int *rowIdx, int *kernel, long long *panel_bytes);

/* Host-only: the tile schedule of the factorised path's contraction kernel (csrc/kernels_contract.cu) for
 * these fragment rows.  counts[3] = { tiles, row slots, strip-table entries (24 per tile) }; the arrays may be
 * NULL (size query), else tileInfo[4 * tiles] = { slotBegin, nSlots, first strip-table entry, nStrips } per
 * tile, slotRow[slots] = fragment row of the slot or -1, strips[tile][warp 0..7][0..2] = b1 | b2 << 8 | n2 << 16
 * (8-row blocks of one fragment, tile-local: block b1 against blocks b2 .. b2 + n2 - 1, b2 >= b1) or -1. */
int qcdmet_contract_schedule(int Nfrag, const int *rowStart, int *counts, int *tileInfo, int *slotRow, int *strips);

const char *qcdmet_last_error(void);
const char *qcdmet_version(void);
int qcdmet_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QCDMET_B200_H */
