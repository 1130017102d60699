// kernels_dmma.cu -- site-factorised FP64 tensor-core (DMMA.8x8x4) Jacobian kernel for sm_100a.
//
// Mathematics (SURVEY.md 8a "mathematical statement", re-associated; exact, no approximation):
//   M_r[a,o]    = sum_v PV[a,v] * Cv[r,v] * W[v,o]            W[v,o] = 1/(e_o - e_v)
//   N_r[a;c,b]  = sum_o M_r[a,o] * Co[c,o] * PO[b,o]          b in the fragment of row a
//   G_p[a,b]    = sum_{(r,c) in H1_p} N_r[a;c,b]
//   J[(f;a,b),p]= 2 (G_p[a,b] + G_p[b,a])
// M_r is shared by every perturbation that has an element in row r, so the dominant GEMM runs
// once per *site* (N of them) instead of once per perturbation (P = N(m+1)/2): 2 S nv no flop per
// site.  This replaces lib/rhf_response.cc:82-108 + src/dmet.py:496-520 for perturbation lists
// whose connected components (perturbations sharing sites) have <= 4 sites and fragments of <= 8
// rows: DMET fits with impurities of <= 4 orbitals, impurity-only or impurity+bath rows.
// Translation-invariant lists (one element per cell) are split into per-cell sub-perturbations
// first and the sub-columns summed afterwards (dD is linear in H1).
//
// One CTA = (component g of <= 4 sites) x (a-tile of 32 fragment rows): a 128 x no output panel
// M[(r,a), o], produced BN = 128 / 96 / 64 columns at a time (Cfg<RF, NJ>, BN = 16 NJ):
//   * B operand: W, pre-tiled in HBM in DMMA fragment order, one TMA bulk copy per stage (16 KB at BN=128)
//   * A operand: (PV[a,:] * Cv[r,:]) built in registers from two TMA-staged panels (4 KB + 512 B)
//   * 8 consumer warps (4 sites x 2 column halves), warp tile 32 x BN/2, 4 NJ DMMA.8x8x4 per k4 step,
//     accumulators in registers (FP64 has no tcgen05/TMEM path: tcgen05.mma has no f64 kind)
//   * 1 producer warp, 4-stage mbarrier full/empty pipeline (cp.async.bulk -> UBLKCP)
//   * epilogue per column tile: accumulators x Co[c,o] x PO[b,o] on the FP64 pipe (operands
//     TMA-prefetched by the producer, each warp runs its epilogue independently so the other warp of
//     the SM sub-partition keeps the DMMA pipe busy), warp-shuffle reduction over the 4 lanes of a row,
//     running sums in shared memory; after the last tile the CTA assembles G_p, symmetrises, and stores
//     its patch of J with coalesced stores.  No atomics: every J element is owned by exactly one CTA.
//   * symmetric mode: when the fragments are exactly the site groups of the perturbation list
//     (impurity-only rows, the standard non-TI DMET fit) M_r[a,o] = M_a[r,o], so only block pairs
//     (g, t) with tile(g) >= t are computed and each CTA also stores the mirrored patch
//     J[rows of g, columns of the blocks in t] from the same N values: half the flops, exact.
//   * multi-GPU: a rank executes a contiguous slice of the block-pair list and stores every J element
//     into all ranks' buffers (CUDA-IPC-mapped peers, NVLink): the all-gather is fused into the kernel.
//   * panel mode (template parameter PANEL, end of this file): the same main loop without the epilogue,
//     M_r[a,o] stored to HBM for the fragments of more than 8 rows that kernels_contract.cu contracts.
#include "common.cuh"
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <numeric>

namespace qcdmet {

namespace {

constexpr int KC = 16;        // virtual orbitals per pipeline stage
constexpr int MIN_STAGES = 4, MAX_STAGES = 8;   // TMA ring depth: as deep as shared memory allows (Cfg::STAGES)
constexpr int RT = 4;         // site slots per component  (= M-warps)
constexpr int AT = 32;        // row slots per a-tile
constexpr int MAXRF = 8;      // max rows per fragment over all variants
constexpr int MAXP = 16;      // max perturbations per component
constexpr int MAXE = 4;       // max elements per perturbation
constexpr int NCONS = 256;    // consumer threads (8 warps)
constexpr int NTHREADS = NCONS + 32;
constexpr int TP = 128;       // max perturbations over the blocks of one tile
constexpr int PV_STAGE = KC * AT;
constexpr int CV_STAGE = KC * RT;

// Item tables: everything a CTA needs to know about a work item (component, tile) lives in two contiguous blobs,
// one per tile and one per component, which the producer copies into shared memory with two TMA bulk copies
// (TABS-deep ring), so the persistent CTA never waits on a dependent global load between two items.
constexpr int TT_QROW = 0;                               // long long [AT]        qoff[f] + a_loc
constexpr int TT_TROW = TT_QROW + AT * 8;                // int [AT]              fragment-row index or -1
constexpr int TT_TP = TT_TROW + AT * 4;                  // int [TP]              perturbations of the blocks in the tile (sym mode)
constexpr int TT_TE = TT_TP + TP * 4;                    // u8 [TP][MAXE]         a_slot | j << 5, 0xFF = none (sym mode)
constexpr int TT_PARTNER = TT_TE + TP * MAXE;            // u8 [AT][RF]           slot of the j-th row of the fragment, AT = none
constexpr int TT_SELF = TT_PARTNER + AT * MAXRF;         // u8 [AT]               a_loc | r_f << 4
constexpr int TILE_TAB_BYTES = TT_SELF + AT;
constexpr int CT_Q = 0;                                  // long long             qoff of the component's own fragment (sym mode)
constexpr int CT_TILE = 8;                               // int                   tile holding that fragment (sym mode)
constexpr int CT_LOC = 12;                               // u8 [RT]               local row index of site slot | r_f << 4 (sym mode)
constexpr int CT_CP = 16;                                // int [MAXP]            perturbation index or -1
constexpr int CT_CENT = CT_CP + MAXP * 4;                // u8 [MAXP][MAXE]       rs | cs << 4, 0xFF = none
constexpr int COMP_TAB_BYTES = CT_CENT + MAXP * MAXE;
constexpr int ITEM_TAB_BYTES = TILE_TAB_BYTES + COMP_TAB_BYTES;
constexpr int TABS = 3;                                  // item-table ring: current item, deferred assembly of the previous one, prefetch
static_assert(TILE_TAB_BYTES % 16 == 0 && COMP_TAB_BYTES % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
constexpr int SMEM_LIMIT = 232448;                       // 227 KB: opt-in dynamic shared memory per CTA on sm_100

// Kernel variant: RF = max rows per fragment (4: impurity-only 2x2 fits; 8: impurity+bath rows of a
// 4-orbital impurity), NJ = 8-column DMMA fragments per warp (column tile BN = 16 NJ: 128 / 96 / 64,
// picked to minimise the padding of the occupied dimension).
template <int RF_, int NJ_>
struct Cfg {
    static constexpr int RF = RF_, NJ = NJ_;
    static constexpr int BN = 16 * NJ;                  // occupied orbitals per output tile
    static constexpr int PSTRIDE = BN + 8;              // POS row stride in doubles (bank-conflict padding)
    static constexpr int W_STAGE = KC * BN;             // doubles
    static constexpr int STAGE_DOUBLES = W_STAGE + PV_STAGE + CV_STAGE;
    static constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int POS_DOUBLES = (AT + 1) * PSTRIDE;
    static constexpr int COS_DOUBLES = RT * BN;
    static constexpr uint32_t EPI_BYTES = (POS_DOUBLES + COS_DOUBLES) * 8;
    static constexpr int NSTRIDE = RT * RF + 2;         // running-sum row stride (conflict-free 16 B slots)
    static constexpr int NS_DOUBLES = 2 * 128 * NSTRIDE;   // running sums of one item: [column half][site slot][row slot][NSTRIDE]
    // barriers: full/empty per stage, epilogue operands full/empty, item tables full/empty, running sums done/free
    static constexpr int BAR_BYTES = (((2 * MAX_STAGES + 2 + 2 * TABS + 4 + 1) * 8 + 15) / 16) * 16;
    static constexpr int AUX_BYTES = BAR_BYTES + TABS * ITEM_TAB_BYTES;
    static constexpr int FIXED1 = (POS_DOUBLES + COS_DOUBLES + NS_DOUBLES) * 8 + AUX_BYTES;
    // two running-sum buffers when they fit next to a minimal ring: the assembly + stores of item k then run inside
    // item k+1, warp by warp, next to the other warps' DMMAs instead of in a CTA-wide bubble
    static constexpr int NSBUF = (MIN_STAGES * (int)STAGE_BYTES + FIXED1 + NS_DOUBLES * 8 <= SMEM_LIMIT) ? 2 : 1;
    static constexpr int FIXED = FIXED1 + (NSBUF - 1) * NS_DOUBLES * 8;
    // ring depth: the rest of the shared memory.  A deep ring lets the two warps of an SM sub-partition run several
    // stages apart, so that one of them is in the DMMA main loop while the other one is in its epilogue.
    static constexpr int STAGES_FIT = (SMEM_LIMIT - FIXED) / (int)STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > MAX_STAGES ? MAX_STAGES : STAGES_FIT;
    static_assert(STAGES >= MIN_STAGES, "shared memory budget");
    static constexpr int SM_STAGE_OFF = 0;
    static constexpr int SM_POS_OFF = SM_STAGE_OFF + STAGES * STAGE_DOUBLES * 8;
    static constexpr int SM_COS_OFF = SM_POS_OFF + POS_DOUBLES * 8;
    static constexpr int SM_NS_OFF = SM_COS_OFF + COS_DOUBLES * 8;
    static constexpr int SM_BAR_OFF = SM_NS_OFF + NSBUF * NS_DOUBLES * 8;
    static constexpr int SM_TAB_OFF = SM_BAR_OFF + BAR_BYTES;
    static constexpr int SMEM_BYTES = SM_TAB_OFF + TABS * ITEM_TAB_BYTES;
    static_assert(SMEM_BYTES <= SMEM_LIMIT, "shared memory budget");
    static_assert(SM_TAB_OFF % 16 == 0 && SM_POS_OFF % 16 == 0 && SM_COS_OFF % 16 == 0, "TMA destinations are 16-byte aligned");
};

struct DmmaArgs {
    int nKc, nOt;
    int nWork;                          // work items of this launch; CTA b executes items b, b + gridDim.x, ...
    int nP, nTP;                        // upper bounds of used perturbation slots per component / per tile (assembly loops)
    int maxE;                           // elements per perturbation (<= MAXE)
    int colBegin, colEnd;
    int sym;                            // symmetric mode: also store the mirrored patch
    long long Q;
    long long jCap;                     // doubles in every destination buffer reachable from column colBegin (bounds checks)
    const double *Wt, *PVt, *CVt, *POt, *COt;
    const int2 *work;                   // (component, tile) per work item
    const unsigned char *tileTab;       // [nTiles][TILE_TAB_BYTES]
    const unsigned char *compTab;       // [nComp][COMP_TAB_BYTES]
    // panel mode (large fragments): no epilogue, the M panels go to HBM as Mout[(4 comp + r)][row][ldm]
    double *Mout;
    int mRows, ldm;
    long long *prof;                    // PROF builds: [gridDim.x][8 warps][8] cycle counters
    // SPLIT builds: a CTA owns a contiguous range of (item, column tile) units; an item cut by a range boundary is finished
    // by the CTA that holds its first tiles, which adds the other CTA's per-tile contributions (ws) in tile order
    double *ws;                         // [gridDim.x][nOt - 1][NS_DOUBLES]
    unsigned int *flags;                // [gridDim.x] consumer warps of range r that have written their share of ws[r]
    int nDest;                          // destinations of every Jacobian store (1, or the ranks of a node)
    double *J[QCDMET_MAX_PEERS];        // J[0..nDest): own buffer and peer-mapped buffers (NVLink P2P stores)
};

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// operand tiling (once per set of orbitals)
// ------------------------------------------------------------------------------------------
// Wt[(ot*nKc + kc)][k4][njp][lane][2] = W[kc*KC + k4*4 + (lane&3)][ot*BN + (2 njp + e)*8 + (lane>>2)]
__global__ void build_wt_kernel(const double *__restrict__ e, int no, int nv, int nKc, int BN, long long total,
                                double *__restrict__ Wt)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int wstage = KC * BN, perK4 = 4 * BN;      // doubles per stage / per k4 step (BN/16 pairs x 32 lanes x 2)
    const int within = (int)(idx % wstage);
    const long long blk = idx / wstage;
    const int ot = (int)(blk / nKc), kc = (int)(blk % nKc);
    const int k4 = within / perK4, r4 = within % perK4;
    const int e2 = r4 & 1, lane = (r4 >> 1) & 31, njp = r4 >> 6;
    const int k = kc * KC + k4 * 4 + (lane & 3);
    const int o = ot * BN + (njp * 2 + e2) * 8 + (lane >> 2);
    Wt[idx] = (k < nv && o < no) ? 1.0 / (e[o] - e[no + k]) : 0.0;
}
// PVt[(tile*nKc + kc)][k4][mi>>1][lane][mi&1] = PV[tileRow[tile][mi*8 + (lane>>2)]][kc*KC + k4*4 + (lane&3)]
// (two conflict-free LDS.128 per k4 step: a quarter warp reads 128 contiguous bytes)
__global__ void build_pvt_kernel(const double *__restrict__ PV, int ldv, int nv, int nKc,
                                 const int *__restrict__ tileRow, long long total, double *__restrict__ PVt)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int within = (int)(idx % PV_STAGE);
    const long long blk = idx / PV_STAGE;
    const int tile = (int)(blk / nKc), kc = (int)(blk % nKc);
    const int lane = (within >> 1) & 31, mi = (within & 1) | (((within >> 6) & 1) << 1), k4 = within >> 7;
    const int row = tileRow[tile * AT + mi * 8 + (lane >> 2)];
    const int k = kc * KC + k4 * 4 + (lane & 3);
    PVt[idx] = (row >= 0 && k < nv) ? PV[(size_t)row * ldv + k] : 0.0;
}
// CVt[(comp*nKc + kc)][r][kk] = CvT[compSites[comp][r]][kc*KC + kk]
__global__ void build_cvt_kernel(const double *__restrict__ CvT, int ldv, int nv, int nKc,
                                 const int *__restrict__ compSites, long long total, double *__restrict__ CVt)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int within = (int)(idx % CV_STAGE);
    const long long blk = idx / CV_STAGE;
    const int comp = (int)(blk / nKc), kc = (int)(blk % nKc);
    const int kk = within % KC, r = within / KC;
    const int site = compSites[comp * RT + r];
    const int k = kc * KC + kk;
    CVt[idx] = (site >= 0 && k < nv) ? CvT[(size_t)site * ldv + k] : 0.0;
}

// POt[(tile*nOt + ot)][slot 0..AT][PSTRIDE] = PO[tileRow[tile][slot]][ot*BN + o]  (slot AT: zero row)
__global__ void build_pot_kernel(const double *__restrict__ PO, int ldo, int no, int nOt, int BN,
                                 const int *__restrict__ tileRow, long long total, double *__restrict__ POt)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int PSTRIDE = BN + 8, POS_DOUBLES = (AT + 1) * PSTRIDE;
    const int within = (int)(idx % POS_DOUBLES);
    const long long blk = idx / POS_DOUBLES;
    const int tile = (int)(blk / nOt), ot = (int)(blk % nOt);
    const int slot = within / PSTRIDE, oo = within % PSTRIDE;
    const int o = ot * BN + oo;
    double val = 0.0;
    if (slot < AT && oo < BN && o < no) {
        const int row = tileRow[tile * AT + slot];
        if (row >= 0) val = PO[(size_t)row * ldo + o];
    }
    POt[idx] = val;
}
// COt[(comp*nOt + ot)][r][BN] = CoT[compSites[comp][r]][ot*BN + o]
__global__ void build_cot_kernel(const double *__restrict__ CoT, int ldo, int no, int nOt, int BN,
                                 const int *__restrict__ compSites, long long total, double *__restrict__ COt)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int COS_DOUBLES = RT * BN;
    const int within = (int)(idx % COS_DOUBLES);
    const long long blk = idx / COS_DOUBLES;
    const int comp = (int)(blk / nOt), ot = (int)(blk % nOt);
    const int r = within / BN, o = ot * BN + within % BN;
    const int site = compSites[comp * RT + r];
    COt[idx] = (site >= 0 && o < no) ? CoT[(size_t)site * ldo + o] : 0.0;
}

// Columns of perturbations without elements (nnz_p = 0) are exactly zero.
struct ZeroArgs { int nDest; double *J[QCDMET_MAX_PEERS]; };
__global__ void zero_columns_kernel(ZeroArgs z, const int *__restrict__ cols, int colBegin, long long Q)
{
    const size_t base = (size_t)Q * (cols[blockIdx.y] - colBegin);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += (long long)gridDim.x * blockDim.x)
        for (int dst = 0; dst < z.nDest; ++dst) z.J[dst][base + q] = 0.0;
}

// Perturbations whose elements fall into several disconnected site groups (translation-invariant
// lists: the same element replicated in every cell, src/dmet.py:139-148) are split into one
// sub-perturbation per group; dD is linear in H1, so J[:,p] = sum over its sub-perturbations
// (fixed order: deterministic).
__global__ void sum_subcolumns_kernel(const double *__restrict__ Jsub, const int *__restrict__ subStart,
                                      int colBegin, int subBegin, long long Q, double *__restrict__ J)
{
    const int p = colBegin + blockIdx.y;
    const int sb = subStart[p], se = subStart[p + 1];
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int k = sb; k < se; ++k) acc += Jsub[(size_t)Q * (k - subBegin) + q];
        J[(size_t)Q * blockIdx.y + q] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
// Persistent: one CTA per SM walks the work list with stride gridDim.x.  The producer warp keeps the TMA ring
// running across work items (the first stages, the item tables and the epilogue operands of the next item are
// in flight while the consumers finish the current one), and with two running-sum buffers the final assembly
// + stores of item k are executed by every consumer warp on its own inside item k+1.
// PANEL = true: no epilogue, the M panels go to HBM (factorised path for large fragments, end of this file)
// PROF = true (tuning builds only, $QCDMET_B200_PROFILE): every consumer warp adds up the clock cycles it spends
// waiting for stages / in the stage bodies / waiting for epilogue operands / in the epilogue / in the assembly.
// SPLIT = true: tile-granular balancing for sizes where nWork is not a multiple of the SM count (see launch_jacobian_dmma).
template <class C, bool PANEL = false, bool PROF = false, bool SPLIT = false>
__global__ void __launch_bounds__(NTHREADS, 1) jac_dmma_kernel(const DmmaArgs g)
{
    static_assert(!(PANEL && SPLIT), "panel mode has no running sums to split");
    constexpr int RF = C::RF, NJ = C::NJ, BN = C::BN, PSTRIDE = C::PSTRIDE, NSTRIDE = C::NSTRIDE;
    constexpr int W_STAGE = C::W_STAGE, STAGE_DOUBLES = C::STAGE_DOUBLES;
    constexpr int POS_DOUBLES = C::POS_DOUBLES, COS_DOUBLES = C::COS_DOUBLES;
    constexpr int NSBUF = C::NSBUF, NS_DOUBLES = C::NS_DOUBLES, STAGES = C::STAGES;
    extern __shared__ __align__(128) unsigned char smem[];
    double *stageBase = reinterpret_cast<double *>(smem + C::SM_STAGE_OFF);
    double *POS = reinterpret_cast<double *>(smem + C::SM_POS_OFF);
    double *CoS = reinterpret_cast<double *>(smem + C::SM_COS_OFF);
    double *Ns = reinterpret_cast<double *>(smem + C::SM_NS_OFF);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::SM_BAR_OFF);
    unsigned char *tabs = smem + C::SM_TAB_OFF;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nKc = g.nKc, nOt = g.nOt, nWork = g.nWork;
    const uint32_t barBase = smem_u32(bars);
    const uint32_t epiFull = barBase + 8 * (2 * STAGES), epiEmpty = epiFull + 8;
    const uint32_t tabFull0 = epiEmpty + 8, tabEmpty0 = tabFull0 + 8 * TABS;
    const uint32_t nsDone0 = tabEmpty0 + 8 * TABS, nsFree0 = nsDone0 + 8 * 2;

    // ---- set-up: barriers
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(barBase + 8 * s, 1);                 // full: producer's expect_tx arrival
            mbar_init(barBase + 8 * (STAGES + s), 8);      // empty: one arrival per consumer warp
        }
        mbar_init(epiFull, 1);
        mbar_init(epiEmpty, 8);
        for (int t = 0; t < TABS; ++t) { mbar_init(tabFull0 + 8 * t, 1); mbar_init(tabEmpty0 + 8 * t, 8); }
        for (int b = 0; b < 2; ++b) { mbar_init(nsDone0 + 8 * b, 8); mbar_init(nsFree0 + 8 * b, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        // ================= producer warp: one elected lane drives the TMA pipeline =================
        if (lane == 0) {
            const int epiAt = nKc > 3 ? nKc - 3 : 0;      // prefetch the epilogue operands near the end of a tile
            int rs = 0, rph = 0, T = 0, k = 0;           // ring slot and phase of the next stage
            int2 wkNext = blockIdx.x < nWork ? g.work[blockIdx.x] : make_int2(0, 0);
            // SPLIT: units [u, uEnd) of range r = gridDim.x - 1 - blockIdx.x (a CTA only ever waits for a lower-numbered CTA)
            long long u = 0, uEnd = 0;
            if constexpr (SPLIT) {
                const long long units = (long long)nWork * nOt, r = gridDim.x - 1 - blockIdx.x;
                u = units * r / gridDim.x;
                uEnd = units * (r + 1) / gridDim.x;
            }
            for (int w = blockIdx.x; SPLIT ? u < uEnd : w < nWork; w += gridDim.x, ++k) {
                int comp, tile, ota = 0, otb = nOt;
                if constexpr (SPLIT) {
                    const int item = (int)(u / nOt);
                    ota = (int)(u % nOt);
                    otb = (int)min((long long)nOt, ota + (uEnd - u));
                    u += otb - ota;
                    const int2 wk = g.work[item];
                    comp = wk.x; tile = wk.y;
                } else {
                    comp = wkNext.x; tile = wkNext.y;
                    if (w + (int)gridDim.x < nWork) wkNext = g.work[w + gridDim.x];
                }
                // item tables
                const int tb = k % TABS;
                mbar_wait(tabEmpty0 + 8 * tb, ((k / TABS) & 1) ^ 1);
                mbar_expect_tx(tabFull0 + 8 * tb, ITEM_TAB_BYTES);
                unsigned char *tdst = tabs + tb * ITEM_TAB_BYTES;
                tma_bulk_g2s(smem_u32(tdst), g.tileTab + (size_t)tile * TILE_TAB_BYTES, TILE_TAB_BYTES, tabFull0 + 8 * tb);
                tma_bulk_g2s(smem_u32(tdst + TILE_TAB_BYTES), g.compTab + (size_t)comp * COMP_TAB_BYTES, COMP_TAB_BYTES,
                             tabFull0 + 8 * tb);
                const double *pvsrc = g.PVt + (size_t)tile * nKc * PV_STAGE;
                const double *cvsrc = g.CVt + (size_t)comp * nKc * CV_STAGE;
                const double *posrc = g.POt + (size_t)tile * nOt * POS_DOUBLES;
                const double *cosrc = g.COt + (size_t)comp * nOt * COS_DOUBLES;
                for (int ot = ota; ot < otb; ++ot, ++T) {
                    const double *wsrc = g.Wt + (size_t)ot * nKc * W_STAGE;
                    for (int kc = 0; kc < nKc; ++kc) {
                        if (!PANEL && kc == epiAt) {
                            mbar_wait(epiEmpty, (T & 1) ^ 1);       // every warp is done with the previous tile's operands
                            mbar_expect_tx(epiFull, C::EPI_BYTES);
                            tma_bulk_g2s(smem_u32(POS), posrc + (size_t)ot * POS_DOUBLES, POS_DOUBLES * 8, epiFull);
                            tma_bulk_g2s(smem_u32(CoS), cosrc + (size_t)ot * COS_DOUBLES, COS_DOUBLES * 8, epiFull);
                        }
                        const uint32_t full = barBase + 8 * rs, empty = barBase + 8 * (STAGES + rs);
                        mbar_wait(empty, rph ^ 1);
                        mbar_expect_tx(full, C::STAGE_BYTES);
                        double *dst = stageBase + (size_t)rs * STAGE_DOUBLES;
                        tma_bulk_g2s(smem_u32(dst), wsrc + (size_t)kc * W_STAGE, W_STAGE * 8, full);
                        tma_bulk_g2s(smem_u32(dst + W_STAGE), pvsrc + (size_t)kc * PV_STAGE, PV_STAGE * 8, full);
                        tma_bulk_g2s(smem_u32(dst + W_STAGE + PV_STAGE), cvsrc + (size_t)kc * CV_STAGE, CV_STAGE * 8, full);
                        if (++rs == STAGES) { rs = 0; rph ^= 1; }
                    }
                }
            }
        }
        return;
    }

    // ================= consumer warps =================
    const int warp_m = warp >> 1;     // site slot r
    const int warp_n = warp & 1;      // column half: NJ fragments of 8 columns
    const int g4 = lane >> 2, t4 = lane & 3;

    // Assemble G_p of work item j from its running sums, symmetrise, store the item's patches of J.  Every
    // consumer thread owns the same share of the index space whenever it gets here: no CTA-wide barrier, only
    // "all eight warps have added their last column tile" (nsDone).
    auto assemble = [&](int j, int jn, int tile) {       // j: part (item tables), jn: its running-sum use; equal unless SPLIT
        const int tbj = j % TABS, nbj = jn % NSBUF;
        mbar_wait(nsDone0 + 8 * nbj, (jn / NSBUF) & 1);
        const unsigned char *tab = tabs + tbj * ITEM_TAB_BYTES;
        const long long *qrow = reinterpret_cast<const long long *>(tab + TT_QROW);
        const int *trow = reinterpret_cast<const int *>(tab + TT_TROW);
        const unsigned char *partner = tab + TT_PARTNER, *selfi = tab + TT_SELF;
        const unsigned char *ctab = tab + TILE_TAB_BYTES;
        const int *cp = reinterpret_cast<const int *>(ctab + CT_CP);
        const unsigned char *cent = ctab + CT_CENT;
        const double *Ns0 = Ns + nbj * NS_DOUBLES, *Ns1 = Ns0 + 128 * NSTRIDE;
        // Two forms of the same arithmetic.  ptxas schedules the DMMA main loop of a variant differently depending on
        // the code around it: measured on one B200 (profiles/r2_ab_fused_assembly.txt), the branchy form below keeps
        // Cfg<4,6> 2.9 % faster (N = 576 and N = 4000), the branch-free form keeps Cfg<4,8> 1.8 % faster (N = 1000, 2000).
        if constexpr (RF == 4 && NJ == 6) {
            // patch 1: rows of the tile x perturbations of the component
            for (int idx = tid; idx < MAXP * AT * RF; idx += NCONS) {
                const int pl = idx / (AT * RF), rem = idx % (AT * RF);
                const int a_slot = rem / RF, jj = rem % RF;
                const int p = cp[pl];
                if (p < g.colBegin || p >= g.colEnd) continue;
                const int self = selfi[a_slot];
                const int a_loc = self & 15, rf = self >> 4;
                if (trow[a_slot] < 0 || jj >= rf) continue;
                const int b_slot = partner[a_slot * RF + jj];
                double gab = 0.0, gba = 0.0;
    #pragma unroll
                for (int e = 0; e < MAXE; ++e) {
                    const unsigned char ent = cent[pl * MAXE + e];
                    if (ent == 0xFF) continue;
                    const int rs = ent & 15, cs = ent >> 4;
                    const int i1 = (rs * AT + a_slot) * NSTRIDE + cs * RF + jj;
                    const int i2 = (rs * AT + b_slot) * NSTRIDE + cs * RF + a_loc;
                    gab += Ns0[i1] + Ns1[i1];
                    gba += Ns0[i2] + Ns1[i2];
                }
                const size_t off = (size_t)g.Q * (p - g.colBegin) + qrow[a_slot] + (long long)rf * jj;
                const double val = 2.0 * (gab + gba);
                QC_DEV_ASSERT(off < (size_t)g.jCap);
                for (int dst = 0; dst < g.nDest; ++dst) g.J[dst][off] = val;
            }
            // patch 2 (symmetric mode, off-diagonal block pairs): rows of the component's own fragment x
            // perturbations of the blocks in the tile, from M_a[r,o] = M_r[a,o]:
            //   G_p'[r,c] = sum_{(a,b') in p'} N_r[a; c, b']
            if (g.sym && *reinterpret_cast<const int *>(ctab + CT_TILE) != tile) {
                const long long q0 = *reinterpret_cast<const long long *>(ctab + CT_Q);
                const unsigned char *loc = ctab + CT_LOC;
                const int *tp = reinterpret_cast<const int *>(tab + TT_TP);
                const unsigned char *te = tab + TT_TE;
                for (int idx = tid; idx < TP * RT * RT; idx += NCONS) {
                    const int pp = idx / (RT * RT), rem = idx % (RT * RT);
                    const int cs = rem / RT, rs = rem % RT;          // site slots of the component: row r, column c
                    const int p = tp[pp];
                    if (p < g.colBegin || p >= g.colEnd) continue;
                    const int lr = loc[rs], lc = loc[cs];
                    if (lr == 0xFF || lc == 0xFF) continue;
                    const int rf = lr >> 4;
                    double grc = 0.0, gcr = 0.0;
    #pragma unroll
                    for (int e = 0; e < MAXE; ++e) {
                        const unsigned char ent = te[pp * MAXE + e];
                        if (ent == 0xFF) continue;
                        const int a_slot = ent & 31, jj = ent >> 5;
                        const int i1 = (rs * AT + a_slot) * NSTRIDE + cs * RF + jj;
                        const int i2 = (cs * AT + a_slot) * NSTRIDE + rs * RF + jj;
                        grc += Ns0[i1] + Ns1[i1];
                        gcr += Ns0[i2] + Ns1[i2];
                    }
                    const size_t off = (size_t)g.Q * (p - g.colBegin) + q0 + (lr & 15) + (long long)rf * (lc & 15);
                    const double val = 2.0 * (grc + gcr);
                    QC_DEV_ASSERT(off < (size_t)g.jCap);
                    for (int dst = 0; dst < g.nDest; ++dst) g.J[dst][off] = val;
                }
            }
        } else {
            // patch 1: rows of the tile x perturbations of the component.  Branch-free: elements a perturbation does not
            // have enter with weight 0 (every running sum is finite), only the store is predicated.
            const int maxE = g.maxE;
            for (int idx = tid; idx < g.nP * AT * RF; idx += NCONS) {
                const int pl = idx / (AT * RF), rem = idx % (AT * RF);
                const int a_slot = rem / RF, jj = rem % RF;
                const int p = cp[pl];
                const int self = selfi[a_slot];
                const int a_loc = self & 15, rf = self >> 4;
                const int b_slot = partner[a_slot * RF + jj] & (AT - 1);       // "none" (AT) only where nothing is stored
                const int o1 = a_slot * NSTRIDE + jj, o2 = b_slot * NSTRIDE + a_loc;
                double gab = 0.0, gba = 0.0;
                for (int e = 0; e < maxE; ++e) {
                    const unsigned char ent = cent[pl * MAXE + e];
                    const double wgt = ent == 0xFF ? 0.0 : 1.0;
                    const int rs = ent & 3, cs = (ent >> 4) & 3;
                    const int base = rs * AT * NSTRIDE + cs * RF;
                    gab = fma(wgt, Ns0[base + o1] + Ns1[base + o1], gab);
                    gba = fma(wgt, Ns0[base + o2] + Ns1[base + o2], gba);
                }
                if (p >= g.colBegin && p < g.colEnd && trow[a_slot] >= 0 && jj < rf) {
                    const size_t off = (size_t)g.Q * (p - g.colBegin) + qrow[a_slot] + (long long)rf * jj;
                    const double val = 2.0 * (gab + gba);
                    QC_DEV_ASSERT(off < (size_t)g.jCap);
                    for (int dst = 0; dst < g.nDest; ++dst) g.J[dst][off] = val;
                }
            }
            // patch 2 (symmetric mode, off-diagonal block pairs): rows of the component's own fragment x
            // perturbations of the blocks in the tile, from M_a[r,o] = M_r[a,o]:
            //   G_p'[r,c] = sum_{(a,b') in p'} N_r[a; c, b']
            if (g.sym && *reinterpret_cast<const int *>(ctab + CT_TILE) != tile) {
                const long long q0 = *reinterpret_cast<const long long *>(ctab + CT_Q);
                const unsigned char *loc = ctab + CT_LOC;
                const int *tp = reinterpret_cast<const int *>(tab + TT_TP);
                const unsigned char *te = tab + TT_TE;
                for (int idx = tid; idx < g.nTP * RT * RT; idx += NCONS) {
                    const int pp = idx / (RT * RT), rem = idx % (RT * RT);
                    const int cs = rem / RT, rs = rem % RT;          // site slots of the component: row r, column c
                    const int p = tp[pp];
                    const int lr = loc[rs], lc = loc[cs];
                    const int rf = lr >> 4;
                    const int o1 = rs * AT * NSTRIDE + cs * RF, o2 = cs * AT * NSTRIDE + rs * RF;
                    double grc = 0.0, gcr = 0.0;
                    for (int e = 0; e < maxE; ++e) {
                        const unsigned char ent = te[pp * MAXE + e];
                        const double wgt = ent == 0xFF ? 0.0 : 1.0;
                        const int base = (ent & 31) * NSTRIDE + ((ent >> 5) & 3);
                        grc = fma(wgt, Ns0[base + o1] + Ns1[base + o1], grc);
                        gcr = fma(wgt, Ns0[base + o2] + Ns1[base + o2], gcr);
                    }
                    if (p >= g.colBegin && p < g.colEnd && lr != 0xFF && lc != 0xFF) {
                        const size_t off = (size_t)g.Q * (p - g.colBegin) + q0 + (lr & 15) + (long long)rf * (lc & 15);
                        const double val = 2.0 * (grc + gcr);
                        QC_DEV_ASSERT(off < (size_t)g.jCap);
                        for (int dst = 0; dst < g.nDest; ++dst) g.J[dst][off] = val;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(nsFree0 + 8 * nbj);          // this warp no longer reads the running sums of item j
            mbar_arrive(tabEmpty0 + 8 * tbj);        // ... nor its tables
        }
    };

    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tPrev = PROF ? clock64() : 0;
    auto lap = [&](int slot) {
        if constexpr (PROF) { const long long t = clock64(); pc[slot] += t - tPrev; tPrev = t; }
    };
    double acc[4][NJ][2];
    int rs = 0, rph = 0, T = 0, k = 0, prevTile = 0;
    // SPLIT: kn counts the parts that use a running-sum buffer (all but the tail part a range may start with), kPrev /
    // knPrev name the part whose assembly is pending; otherwise kn == k and the names fold away
    int kn = 0, kPrev = -1, knPrev = -1;
    long long u = 0, uEnd = 0;
    const int range = SPLIT ? (int)gridDim.x - 1 - (int)blockIdx.x : 0;
    if constexpr (SPLIT) {
        const long long units = (long long)nWork * nOt;
        u = units * range / gridDim.x;
        uEnd = units * (range + 1) / gridDim.x;
    }
    for (int w = blockIdx.x; SPLIT ? u < uEnd : w < nWork; w += gridDim.x, ++k) {
        int item = w, ota = 0, otb = nOt;
        if constexpr (SPLIT) {
            item = (int)(u / nOt);
            ota = (int)(u % nOt);
            otb = (int)min((long long)nOt, ota + (uEnd - u));
            u += otb - ota;
        }
        const bool tail = SPLIT && ota > 0;                  // the item's first tiles belong to the previous range's CTA
        const bool head = SPLIT && ota == 0 && otb < nOt;    // the item's last tiles come from the next range's CTA
        if constexpr (!SPLIT) kn = k;
        const int2 wk = g.work[item];
        const int tb = k % TABS, nb = kn % NSBUF;
        const unsigned char *tab = tabs + tb * ITEM_TAB_BYTES;
        double *NsK = Ns + nb * NS_DOUBLES;
        for (int ot = ota; ot < otb; ++ot, ++T) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int nj = 0; nj < NJ; ++nj) { acc[mi][nj][0] = 0.0; acc[mi][nj][1] = 0.0; }

            for (int kc = 0; kc < nKc; ++kc) {
                lap(7);
                mbar_wait(barBase + 8 * rs, rph);
                lap(0);
                const double *Ws = stageBase + (size_t)rs * STAGE_DOUBLES;
                const double *PVs = Ws + W_STAGE;
                const double *CVs = PVs + PV_STAGE + warp_m * KC;
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    const double2 p01 = *reinterpret_cast<const double2 *>(PVs + ((k4 * 2 + 0) * 32 + lane) * 2);
                    const double2 p23 = *reinterpret_cast<const double2 *>(PVs + ((k4 * 2 + 1) * 32 + lane) * 2);
                    const double cv = CVs[k4 * 4 + t4];
                    double a[4], b[NJ];
                    a[0] = p01.x * cv; a[1] = p01.y * cv; a[2] = p23.x * cv; a[3] = p23.y * cv;
#pragma unroll
                    for (int j = 0; j < NJ / 2; ++j) {
                        const double2 bb = *reinterpret_cast<const double2 *>(
                            Ws + ((k4 * NJ + warp_n * (NJ / 2) + j) * 32 + lane) * 2);
                        b[2 * j] = bb.x; b[2 * j + 1] = bb.y;
                    }
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int nj = 0; nj < NJ; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], a[mi], b[nj]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(barBase + 8 * (STAGES + rs));
                if (++rs == STAGES) { rs = 0; rph ^= 1; }
                lap(1);
            }

            if constexpr (PANEL) {
                // ---- panel mode: store this tile of M_r[a, o] (fragments too large for the on-chip epilogue)
                if (ot == 0) mbar_wait(tabFull0 + 8 * tb, (k / TABS) & 1);
                const int *trow = reinterpret_cast<const int *>(tab + TT_TROW);
                double *mbase = g.Mout + (size_t)(wk.x * RT + warp_m) * g.mRows * g.ldm;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const int row = trow[mi * 8 + g4];
                    if (row < 0) continue;
#pragma unroll
                    for (int nj = 0; nj < NJ; ++nj) {
                        const int o = ot * BN + warp_n * (BN / 2) + nj * 8 + t4 * 2;
                        if (o < g.ldm) {
                            QC_DEV_ASSERT(row < g.mRows);
                            *reinterpret_cast<double2 *>(mbase + (size_t)row * g.ldm + o) =
                                make_double2(acc[mi][nj][0], acc[mi][nj][1]);
                        }
                    }
                }
                continue;
            }
            // ---- epilogue for this column tile (per warp, no CTA-wide barrier)
            if (ot == ota) {
                mbar_wait(tabFull0 + 8 * tb, (k / TABS) & 1);                  // tables of this item
                if (!tail) mbar_wait(nsFree0 + 8 * nb, ((kn / NSBUF) & 1) ^ 1);   // every warp has assembled the buffer's last item
            }
            mbar_wait(epiFull, T & 1);
            lap(2);
            const unsigned char *partner = tab + TT_PARTNER;
            // where this tile's contribution goes: the item's running sums, or (tail part of a split item) its own slot
            // of the workspace, from which the CTA that holds the item's first tiles adds it in tile order
            double *nsTarget = NsK;
            bool startSums = ot == 0;
            if constexpr (SPLIT) {
                if (tail) {
                    nsTarget = g.ws + ((size_t)range * (nOt - 1) + (ot - ota)) * NS_DOUBLES;
                    startSums = true;
                }
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const int a_slot = mi * 8 + g4;
                const double *pc = CoS + warp_n * (BN / 2) + t4 * 2;
                double *nsrow = nsTarget + (warp_n * 128 + warp_m * AT + a_slot) * NSTRIDE + t4 * RF;
#pragma unroll
                for (int jh = 0; jh < RF / 4; ++jh) {            // partners in groups of 4 (register budget)
                    const double *pb[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        pb[j] = POS + (int)partner[a_slot * RF + jh * 4 + j] * PSTRIDE + warp_n * (BN / 2) + t4 * 2;
                    double nacc[RT][4];
#pragma unroll
                    for (int c = 0; c < RT; ++c)
#pragma unroll
                        for (int j = 0; j < 4; ++j) nacc[c][j] = 0.0;
#pragma unroll
                    for (int nj = 0; nj < NJ; ++nj) {
                        double2 po[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) po[j] = *reinterpret_cast<const double2 *>(pb[j] + nj * 8);
                        const double m0 = acc[mi][nj][0], m1 = acc[mi][nj][1];
#pragma unroll
                        for (int c = 0; c < RT; ++c) {
                            const double2 co = *reinterpret_cast<const double2 *>(pc + c * BN + nj * 8);
                            const double t0 = m0 * co.x, t1 = m1 * co.y;
#pragma unroll
                            for (int j = 0; j < 4; ++j) nacc[c][j] = fma(t0, po[j].x, fma(t1, po[j].y, nacc[c][j]));
                        }
                    }
                    // reduce over the 4 lanes that share this row; lane t4 ends up with the sums of site slot c == t4:
                    // butterfly -- exchange the half the partner lane keeps, add the half this lane keeps
                    double half[2][4], keep[4];
#pragma unroll
                    for (int cp2 = 0; cp2 < 2; ++cp2)              // site slots {0,1} -> bit 0, {2,3} -> bit 0
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double mine = (t4 & 1) ? nacc[2 * cp2 + 1][j] : nacc[2 * cp2][j];
                            const double send = (t4 & 1) ? nacc[2 * cp2][j] : nacc[2 * cp2 + 1][j];
                            half[cp2][j] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
                        }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double mine = (t4 & 2) ? half[1][j] : half[0][j];
                        const double send = (t4 & 2) ? half[0][j] : half[1][j];
                        keep[j] = mine + __shfl_xor_sync(0xffffffffu, send, 2);
                    }
                    double2 *dst = reinterpret_cast<double2 *>(nsrow + jh * 4);
                    double2 s01 = make_double2(0.0, 0.0), s23 = make_double2(0.0, 0.0);
                    if (!startSums) { s01 = dst[0]; s23 = dst[1]; }  // the first column tile of an item starts the sums
                    s01.x += keep[0]; s01.y += keep[1]; s23.x += keep[2]; s23.y += keep[3];
                    dst[0] = s01; dst[1] = s23;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(epiEmpty);
            lap(3);
            if constexpr (SPLIT) {
                if (NSBUF == 2 && ot == ota && kPrev >= 0) { assemble(kPrev, knPrev, prevTile); kPrev = -1; lap(4); }
            } else {
                if (NSBUF == 2 && ot == 0 && k > 0) { assemble(k - 1, k - 1, prevTile); lap(4); }   // acc is dead here: no register pressure
            }
        }
        if constexpr (PANEL) {
            __syncwarp();
            if (lane == 0) mbar_arrive(tabEmpty0 + 8 * tb);
            continue;
        }
        if constexpr (SPLIT) {
            if (tail) {
                // this warp's share of the tile contributions is in the workspace: publish, release the item tables
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    atomicAdd(g.flags + range, 1u);
                    mbar_arrive(tabEmpty0 + 8 * tb);
                }
                continue;
            }
            if (head) {
                // add the tiles the next range's CTA computed (it starts with them: they are long done), in tile order --
                // the same association as an unsplit item
                if (lane == 0) {
                    unsigned int seen = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(g.flags + range + 1) : "memory");
                        if (seen < 8u) __nanosleep(200);
                    } while (seen < 8u);
                }
                __syncwarp();
                const int base = (warp_n * 128 + warp_m * AT) * NSTRIDE;
                for (int t = 0; t < nOt - otb; ++t) {
                    const double *src = g.ws + ((size_t)(range + 1) * (nOt - 1) + t) * NS_DOUBLES + base;
                    for (int i = lane; i < AT * NSTRIDE; i += 32)
                        if (i % NSTRIDE < RT * RF) NsK[base + i] += __ldcg(src + i);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(nsDone0 + 8 * nb);       // this warp has added its last column tile of item k
        if constexpr (SPLIT) {
            if (NSBUF == 1) { assemble(k, kn, wk.y); lap(4); }
            else { kPrev = k; knPrev = kn; }
            ++kn;
        } else {
            if (NSBUF == 1) { assemble(k, k, wk.y); lap(4); }
        }
        prevTile = wk.y;
    }
    if constexpr (!PANEL) {
        if constexpr (SPLIT) {
            if (NSBUF == 2 && kPrev >= 0) { assemble(kPrev, knPrev, prevTile); lap(4); }
        } else {
            if (NSBUF == 2 && k > 0) { assemble(k - 1, k - 1, prevTile); lap(4); }
        }
    }
    if constexpr (PROF)
        if (lane == 0 && g.prof)
            for (int i = 0; i < 8; ++i) g.prof[((size_t)blockIdx.x * 8 + warp) * 8 + i] = pc[i];
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side: schedule (components, a-tiles), operand buffers, launch
// ------------------------------------------------------------------------------------------
struct DmmaPlanData {
    int N = 0, no = 0, nv = 0, P = 0, S = 0;
    int nComp = 0, nTiles = 0, nKc = 0, nOt = 0;
    int rf = 4, nj = 8;                 // kernel variant: rows per fragment (4|8), 8-column fragments per warp (8|6|4)
    int bn() const { return 16 * nj; }
    size_t wStage() const { return (size_t)KC * bn(); }
    size_t posDoubles() const { return (size_t)(AT + 1) * (bn() + 8); }
    size_t cosDoubles() const { return (size_t)RT * bn(); }
    bool expanded = false;              // perturbations were split into connected sub-perturbations
    int P2 = 0;                         // number of sub-perturbations (internal columns)
    std::vector<int> subStart;          // [P+1] sub-perturbations of p: [subStart[p], subStart[p+1])
    int *d_subStart = nullptr;
    double *d_Jsub = nullptr;           // scratch (Q x internal columns) when expanded
    size_t JsubCapacity = 0;
    std::vector<int> emptyP;            // perturbations without elements: zero columns
    int *d_emptyCols = nullptr;
    std::vector<int> emptyInRange;
    bool symOk = false;                 // fragments == site groups of the perturbation list
    std::vector<int> compSites, compP, compNumSites, compTile;
    std::vector<long long> compQ;
    std::vector<unsigned char> compEnt, compLoc;
    std::vector<int> tileRow, tileNumRows, tileP;
    std::vector<unsigned char> tilePartner, tileSelf, tileEnt;
    std::vector<long long> tileQ;
    std::vector<int2> work;
    int launchBegin = -1, launchEnd = -1, launchRank = -1, launchRanks = -1;
    bool launchSym = false;
    double launchFlops = 0.0;
    std::vector<unsigned char> tileTab, compTab;    // item tables in the kernel's blob layout (pack_item_tables)
    int *d_compSites = nullptr, *d_tileRow = nullptr;
    unsigned char *d_tileTab = nullptr, *d_compTab = nullptr;
    int numSMs = 0;                     // persistent grid: one CTA per SM of the plan's device
    int nP = 0, nTP = 0, maxE = 0;      // used perturbation slots per component / per tile (max), elements per perturbation (max)
    double *d_ws = nullptr;             // tile-granular split (SPLIT kernel): per-CTA workspace of tile contributions
    size_t wsCapacity = 0;
    unsigned int *d_flags = nullptr;    // [numSMs]
    bool launchSplit = false;
    int2 *d_work = nullptr;
    size_t workCapacity = 0;
    double *d_Wt = nullptr, *d_PVt = nullptr, *d_CVt = nullptr, *d_POt = nullptr, *d_COt = nullptr;
    bool attrSet = false;
};

namespace {

struct Schedule {
    bool ok = false;
    std::string why;
    DmmaPlanData d;
};

Schedule build_schedule(int N, int no, int Porig, int F, const int *h1start0, const int *h1row0, const int *h1col0,
                        const int *rowStart, const int *rowIdx)
{
    Schedule sc;
    DmmaPlanData &d = sc.d;
    d.N = N; d.no = no; d.nv = N - no; d.P = Porig; d.S = F > 0 ? rowStart[F] : 0;
    if (Porig == 0 || d.S == 0) { sc.why = "empty problem"; return sc; }
    // split every perturbation into the connected groups of its own elements (sites linked by an element)
    std::vector<int> xs(1, 0), xr, xc;
    d.subStart.assign(1, 0);
    {
        std::map<int, int> par;
        std::function<int(int)> fnd = [&](int x) { while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; } return x; };
        for (int p = 0; p < Porig; ++p) {
            par.clear();
            for (int k = h1start0[p]; k < h1start0[p + 1]; ++k) {
                if (!par.count(h1row0[k])) par[h1row0[k]] = h1row0[k];
                if (!par.count(h1col0[k])) par[h1col0[k]] = h1col0[k];
            }
            for (int k = h1start0[p]; k < h1start0[p + 1]; ++k) {
                const int ra = fnd(h1row0[k]), rb = fnd(h1col0[k]);
                if (ra != rb) par[rb] = ra;
            }
            std::map<int, std::vector<int>> groups;      // root -> element indices, in original order
            std::vector<int> order;
            for (int k = h1start0[p]; k < h1start0[p + 1]; ++k) {
                const int r = fnd(h1row0[k]);
                if (!groups.count(r)) order.push_back(r);
                groups[r].push_back(k);
            }
            for (int r : order) {
                for (int k : groups[r]) { xr.push_back(h1row0[k]); xc.push_back(h1col0[k]); }
                xs.push_back((int)xr.size());
            }
            d.subStart.push_back((int)xs.size() - 1);
            if (order.size() > 1) d.expanded = true;
        }
    }
    const int P = d.expanded ? (int)xs.size() - 1 : Porig;
    const int *h1start = d.expanded ? xs.data() : h1start0;
    const int *h1row = d.expanded ? xr.data() : h1row0;
    const int *h1col = d.expanded ? xc.data() : h1col0;
    d.P2 = P;
    // fragments: <= 8 rows each, packed whole into tiles of AT row slots
    std::vector<long long> qoff(F + 1, 0);
    int maxRf = 0;
    for (int f = 0; f < F; ++f) {
        const int rf = rowStart[f + 1] - rowStart[f];
        if (rf > MAXRF) { sc.why = "a fragment has more than 8 rows"; return sc; }
        maxRf = std::max(maxRf, rf);
        qoff[f + 1] = qoff[f] + (long long)rf * rf;
    }
    const int RF = maxRf <= 4 ? 4 : 8;
    d.rf = RF;
    // column tile: the width (128 / 96 / 64) that pads the occupied dimension least
    {
        int best = 8;
        long long bestPad = -1;
        for (int nj : {8, 6, 4}) {
            const long long bn = 16 * nj, pad = (no + bn - 1) / bn * bn;
            if (bestPad < 0 || pad < bestPad) { bestPad = pad; best = nj; }
        }
        d.nj = best;
    }
    std::vector<int> fragTile(F, -1), fragSlot(F, -1);
    int slot = AT;   // forces a new tile at the first fragment
    for (int f = 0; f < F; ++f) {
        const int rs = rowStart[f], rf = rowStart[f + 1] - rs;
        if (rf == 0) continue;
        if (slot + rf > AT) {
            d.tileRow.insert(d.tileRow.end(), AT, -1);
            d.tilePartner.insert(d.tilePartner.end(), AT * RF, (unsigned char)AT);
            d.tileSelf.insert(d.tileSelf.end(), AT, 0);
            d.tileQ.insert(d.tileQ.end(), AT, 0);
            d.tileNumRows.push_back(0);
            ++d.nTiles;
            slot = 0;
        }
        const int base = (d.nTiles - 1) * AT;
        fragTile[f] = d.nTiles - 1;
        fragSlot[f] = slot;
        for (int al = 0; al < rf; ++al) {
            d.tileRow[base + slot + al] = rs + al;
            d.tileSelf[base + slot + al] = (unsigned char)(al | (rf << 4));
            d.tileQ[base + slot + al] = qoff[f] + al;
            for (int j = 0; j < rf; ++j) d.tilePartner[(size_t)(base + slot + al) * RF + j] = (unsigned char)(slot + j);
        }
        d.tileNumRows.back() += rf;
        slot += rf;
    }
    // components: perturbations connected through shared sites (union-find over sites)
    std::vector<int> parent(N);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    std::vector<char> used(N, 0);
    for (int p = 0; p < P; ++p) {
        const int b = h1start[p], e = h1start[p + 1];
        if (e == b) { d.emptyP.push_back(p); continue; }
        if (e - b > MAXE) { sc.why = "a perturbation has more than 4 elements (translation-invariant list?)"; return sc; }
        for (int k = b; k < e; ++k) {
            used[h1row[k]] = used[h1col[k]] = 1;
            const int ra = find(h1row[b]);
            const int rb = find(h1row[k]);
            if (ra != rb) parent[rb] = ra;
            const int rc = find(h1col[k]);
            const int rd = find(h1row[b]);
            if (rc != rd) parent[rc] = rd;
        }
    }
    std::map<int, int> compOfRoot;
    std::vector<std::vector<int>> sites;
    for (int i = 0; i < N; ++i) {
        if (!used[i]) continue;
        const int r = find(i);
        auto itc = compOfRoot.find(r);
        int c;
        if (itc == compOfRoot.end()) { c = (int)sites.size(); compOfRoot[r] = c; sites.emplace_back(); }
        else c = itc->second;
        sites[c].push_back(i);
        if ((int)sites[c].size() > RT) { sc.why = "perturbations couple more than 4 sites"; return sc; }
    }
    d.nComp = (int)sites.size();
    d.compSites.assign((size_t)d.nComp * RT, -1);
    d.compP.assign((size_t)d.nComp * MAXP, -1);
    d.compEnt.assign((size_t)d.nComp * MAXP * MAXE, 0xFF);
    d.compNumSites.assign(d.nComp, 0);
    std::vector<int> np(d.nComp, 0);
    for (int c = 0; c < d.nComp; ++c) {
        d.compNumSites[c] = (int)sites[c].size();
        for (size_t k = 0; k < sites[c].size(); ++k) d.compSites[(size_t)c * RT + k] = sites[c][k];
    }
    std::vector<int> compOfP(P, -1);
    for (int p = 0; p < P; ++p) {
        const int b = h1start[p], e = h1start[p + 1];
        if (e == b) continue;
        const int c = compOfRoot[find(h1row[b])];
        if (np[c] >= MAXP) { sc.why = "more than 16 perturbations on one group of sites"; return sc; }
        const int pl = np[c]++;
        compOfP[p] = c;
        d.compP[(size_t)c * MAXP + pl] = p;
        for (int k = b; k < e; ++k) {
            int rsl = -1, csl = -1;
            for (int q = 0; q < RT; ++q) {
                if (d.compSites[(size_t)c * RT + q] == h1row[k]) rsl = q;
                if (d.compSites[(size_t)c * RT + q] == h1col[k]) csl = q;
            }
            d.compEnt[((size_t)c * MAXP + pl) * MAXE + (k - b)] = (unsigned char)(rsl | (csl << 4));
        }
    }
    d.nKc = (d.nv + KC - 1) / KC;
    d.nOt = (d.no + d.bn() - 1) / d.bn();
    for (int c = 0; c < d.nComp; ++c) d.nP = std::max(d.nP, np[c]);
    for (int p = 0; p < P; ++p) d.maxE = std::max(d.maxE, h1start[p + 1] - h1start[p]);

    // symmetric mode: impurity-only rows, and the row sites of every non-empty fragment are exactly the
    // sites of one component (and vice versa).  Then PV/PO rows are C_v/C_o site rows and
    // M_r[a,o] = M_a[r,o].
    d.compTile.assign(d.nComp, -1);
    d.compQ.assign(d.nComp, 0);
    d.compLoc.assign((size_t)d.nComp * RT, 0xFF);
    d.tileP.assign((size_t)d.nTiles * TP, -1);
    d.tileEnt.assign((size_t)d.nTiles * TP * MAXE, 0xFF);
    bool sym = (rowIdx != nullptr) && RF == 4 && !d.expanded;
    if (sym) {
        std::vector<int> siteComp(N, -1), compFrag(d.nComp, -1);
        for (int c = 0; c < d.nComp; ++c)
            for (int s : sites[c]) siteComp[s] = c;
        int matched = 0;
        for (int f = 0; f < F && sym; ++f) {
            const int rs = rowStart[f], rf = rowStart[f + 1] - rs;
            if (rf == 0) continue;
            const int c = siteComp[rowIdx[rs]];
            if (c < 0 || compFrag[c] >= 0 || d.compNumSites[c] != rf) { sym = false; break; }
            for (int al = 0; al < rf && sym; ++al) {
                int q = -1;
                for (int k = 0; k < RT; ++k)
                    if (d.compSites[(size_t)c * RT + k] == rowIdx[rs + al]) q = k;
                if (q < 0 || d.compLoc[(size_t)c * RT + q] != 0xFF) sym = false;
                else d.compLoc[(size_t)c * RT + q] = (unsigned char)(al | (rf << 4));
            }
            compFrag[c] = f;
            d.compTile[c] = fragTile[f];
            d.compQ[c] = qoff[f];
            ++matched;
        }
        if (matched != d.nComp) sym = false;
        if (sym) {
            // per tile: the perturbations of its blocks, elements as (row slot of a, local index of b')
            std::vector<int> ntp(d.nTiles, 0);
            for (int p = 0; p < P && sym; ++p) {
                const int c = compOfP[p];
                if (c < 0) continue;
                const int f = compFrag[c], t = fragTile[f];
                if (ntp[t] >= TP) { sym = false; break; }
                const int pp = ntp[t]++;
                d.nTP = std::max(d.nTP, ntp[t]);
                d.tileP[(size_t)t * TP + pp] = p;
                for (int k = h1start[p]; k < h1start[p + 1]; ++k) {
                    int al = -1, bl = -1;
                    for (int x = rowStart[f]; x < rowStart[f + 1]; ++x) {
                        if (rowIdx[x] == h1row[k]) al = x - rowStart[f];
                        if (rowIdx[x] == h1col[k]) bl = x - rowStart[f];
                    }
                    d.tileEnt[((size_t)t * TP + pp) * MAXE + (k - h1start[p])] =
                        (unsigned char)((fragSlot[f] + al) | (bl << 5));
                }
            }
        }
    }
    d.symOk = sym;
    sc.ok = true;
    return sc;
}

// Item tables in the layout the kernel copies into shared memory (TT_* / CT_* offsets).  Tables a schedule does not
// have (panel mode: no perturbations, no symmetric mode) keep their "none" defaults.
void pack_item_tables(DmmaPlanData &d)
{
    const int RF = d.rf;
    d.tileTab.assign((size_t)d.nTiles * TILE_TAB_BYTES, 0);
    for (int t = 0; t < d.nTiles; ++t) {
        unsigned char *b = d.tileTab.data() + (size_t)t * TILE_TAB_BYTES;
        long long *qrow = reinterpret_cast<long long *>(b + TT_QROW);
        int *trow = reinterpret_cast<int *>(b + TT_TROW), *tp = reinterpret_cast<int *>(b + TT_TP);
        for (int a = 0; a < AT; ++a) {
            qrow[a] = d.tileQ.empty() ? 0 : d.tileQ[(size_t)t * AT + a];
            trow[a] = d.tileRow[(size_t)t * AT + a];
            b[TT_SELF + a] = d.tileSelf.empty() ? 0 : d.tileSelf[(size_t)t * AT + a];
            for (int j = 0; j < RF; ++j)
                b[TT_PARTNER + a * RF + j] = d.tilePartner.empty() ? (unsigned char)AT : d.tilePartner[((size_t)t * AT + a) * RF + j];
        }
        for (int k = 0; k < TP; ++k) {
            tp[k] = d.tileP.empty() ? -1 : d.tileP[(size_t)t * TP + k];
            for (int e = 0; e < MAXE; ++e)
                b[TT_TE + k * MAXE + e] = d.tileEnt.empty() ? 0xFF : d.tileEnt[((size_t)t * TP + k) * MAXE + e];
        }
    }
    d.compTab.assign((size_t)d.nComp * COMP_TAB_BYTES, 0);
    for (int c = 0; c < d.nComp; ++c) {
        unsigned char *b = d.compTab.data() + (size_t)c * COMP_TAB_BYTES;
        *reinterpret_cast<long long *>(b + CT_Q) = d.compQ.empty() ? 0 : d.compQ[c];
        *reinterpret_cast<int *>(b + CT_TILE) = d.compTile.empty() ? -1 : d.compTile[c];
        int *cp = reinterpret_cast<int *>(b + CT_CP);
        for (int r = 0; r < RT; ++r) b[CT_LOC + r] = d.compLoc.empty() ? 0xFF : d.compLoc[(size_t)c * RT + r];
        for (int k = 0; k < MAXP; ++k) {
            cp[k] = d.compP.empty() ? -1 : d.compP[(size_t)c * MAXP + k];
            for (int e = 0; e < MAXE; ++e)
                b[CT_CENT + k * MAXE + e] = d.compEnt.empty() ? 0xFF : d.compEnt[((size_t)c * MAXP + k) * MAXE + e];
        }
    }
}

int query_sms(DmmaPlanData &d)
{
    int dev = 0, sms = 0;
    QC_CUDA(cudaGetDevice(&dev));
    QC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    d.numSMs = sms > 0 ? sms : 1;
    return QCDMET_OK;
}

template <typename T>
int upload(T **dst, const std::vector<T> &src)
{
    const size_t n = std::max<size_t>(src.size(), 1);
    if (cudaMalloc((void **)dst, n * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc failed (DMMA schedule, %zu bytes)", n * sizeof(T));
        return QCDMET_E_ALLOC;
    }
    if (!src.empty()) QC_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return QCDMET_OK;
}

}  // namespace

bool dmma_eligible(int N, int P, int F, const int *h1start, const int *h1row, const int *h1col,
                   const int *rowStart, const int *rowIdx, bool, std::string *why)
{
    Schedule sc = build_schedule(N, N / 2 > 0 ? N / 2 : 1, P, F, h1start, h1row, h1col, rowStart, rowIdx);
    if (why) *why = sc.why;
    return sc.ok;
}

int dmma_plan_build(DmmaPlanData **out, int N, int no, int P, int F, const int *h1start, const int *h1row,
                    const int *h1col, const int *rowStart, const int *rowIdx)
{
    *out = nullptr;
    Schedule sc = build_schedule(N, no, P, F, h1start, h1row, h1col, rowStart, rowIdx);
    QC_CHECK(sc.ok, QCDMET_E_ARG, "DMMA kernel not eligible: %s", sc.why.c_str());
    DmmaPlanData *d = new DmmaPlanData(std::move(sc.d));
    pack_item_tables(*d);
    int rc = QCDMET_OK;
    if ((rc = query_sms(*d)) || (rc = upload(&d->d_compSites, d->compSites)) || (rc = upload(&d->d_tileRow, d->tileRow)) ||
        (rc = upload(&d->d_tileTab, d->tileTab)) || (rc = upload(&d->d_compTab, d->compTab)) ||
        (rc = upload(&d->d_subStart, d->subStart))) {
        dmma_plan_free(d);
        return rc;
    }
    const size_t wt = (size_t)d->nOt * d->nKc * d->wStage(), pvt = (size_t)d->nTiles * d->nKc * PV_STAGE,
                 cvt = (size_t)d->nComp * d->nKc * CV_STAGE, pot = (size_t)d->nTiles * d->nOt * d->posDoubles(),
                 cot = (size_t)d->nComp * d->nOt * d->cosDoubles();
    if (cudaMalloc((void **)&d->d_Wt, wt * 8) != cudaSuccess || cudaMalloc((void **)&d->d_PVt, pvt * 8) != cudaSuccess ||
        cudaMalloc((void **)&d->d_CVt, cvt * 8) != cudaSuccess || cudaMalloc((void **)&d->d_POt, pot * 8) != cudaSuccess ||
        cudaMalloc((void **)&d->d_COt, cot * 8) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of the tiled operand panels failed (%zu bytes)", (wt + pvt + cvt + pot + cot) * 8);
        dmma_plan_free(d);
        return QCDMET_E_ALLOC;
    }
    *out = d;
    return QCDMET_OK;
}

void dmma_plan_free(DmmaPlanData *d)
{
    if (!d) return;
    cudaFree(d->d_compSites); cudaFree(d->d_tileRow); cudaFree(d->d_tileTab); cudaFree(d->d_compTab); cudaFree(d->d_work);
    cudaFree(d->d_Wt); cudaFree(d->d_PVt); cudaFree(d->d_CVt); cudaFree(d->d_POt); cudaFree(d->d_COt);
    cudaFree(d->d_emptyCols); cudaFree(d->d_subStart); cudaFree(d->d_Jsub); cudaFree(d->d_ws); cudaFree(d->d_flags);
    cudaGetLastError();
    delete d;
}

int dmma_prepare_orbitals(DmmaPlanData *d, const JacView &v, cudaStream_t st)
{
    const long long wt = (long long)d->nOt * d->nKc * d->wStage(), pvt = (long long)d->nTiles * d->nKc * PV_STAGE,
                    cvt = (long long)d->nComp * d->nKc * CV_STAGE, pot = (long long)d->nTiles * d->nOt * d->posDoubles(),
                    cot = (long long)d->nComp * d->nOt * d->cosDoubles();
    build_wt_kernel<<<(unsigned)((wt + 255) / 256), 256, 0, st>>>(v.e, v.no, v.nv, d->nKc, d->bn(), wt, d->d_Wt);
    QC_CUDA(cudaGetLastError());
    build_pvt_kernel<<<(unsigned)((pvt + 255) / 256), 256, 0, st>>>(v.PV, v.ldv, v.nv, d->nKc, d->d_tileRow, pvt, d->d_PVt);
    QC_CUDA(cudaGetLastError());
    build_cvt_kernel<<<(unsigned)((cvt + 255) / 256), 256, 0, st>>>(v.CvT, v.ldv, v.nv, d->nKc, d->d_compSites, cvt, d->d_CVt);
    QC_CUDA(cudaGetLastError());
    build_pot_kernel<<<(unsigned)((pot + 255) / 256), 256, 0, st>>>(v.PO, v.ldo, v.no, d->nOt, d->bn(), d->d_tileRow, pot, d->d_POt);
    QC_CUDA(cudaGetLastError());
    build_cot_kernel<<<(unsigned)((cot + 255) / 256), 256, 0, st>>>(v.CoT, v.ldo, v.no, d->nOt, d->bn(), d->d_compSites, cot, d->d_COt);
    QC_CUDA(cudaGetLastError());
    return QCDMET_OK;
}

int launch_jacobian_dmma(DmmaPlanData *d, const JacView &v, cudaStream_t st, int *launches, double *flops,
                         cudaEvent_t ev0, cudaEvent_t ev1, bool allowSym, const PeerSet *peers)
{
    *launches = 0;
    *flops = 0.0;
    const int ncols = v.colEnd - v.colBegin;
    if (ncols <= 0) return QCDMET_OK;
    QC_CHECK(!(d->expanded && peers), QCDMET_E_STATE,
             "peer mode is not available for split (translation-invariant) perturbation lists");
    // internal column range: sub-perturbations when the list was split
    const int cb = d->expanded ? d->subStart[v.colBegin] : v.colBegin;
    const int ce = d->expanded ? d->subStart[v.colEnd] : v.colEnd;
    const bool wantSym = allowSym && d->symOk && v.colBegin == 0 && v.colEnd == d->P;
    const int shareRank = peers ? peers->rank : 0, shareN = peers ? peers->nranks : 1;
    if (d->launchBegin != v.colBegin || d->launchEnd != v.colEnd || d->launchSym != wantSym ||
        d->launchRank != shareRank || d->launchRanks != shareN) {
        // work list: (component, tile) pairs; symmetric mode (full column range): only pairs with
        // tile(component) >= tile.
        const bool sym = wantSym;
        std::vector<char> hit(d->nComp, 0);
        for (int c = 0; c < d->nComp; ++c)
            for (int k = 0; k < MAXP && !hit[c]; ++k) {
                const int p = d->compP[(size_t)c * MAXP + k];
                hit[c] = (p >= cb && p < ce);
            }
        // Order: blocks of L2BLK components, tile-major inside a block.  Co-resident CTAs share one PVt/POt
        // panel (tile) and the CVt/COt rows of one component block stay L2-resident (128 x 128 KB at
        // N = 4000) next to the Wt panel while every tile passes over them, so each operand panel comes
        // from DRAM about once per component block instead of once per tile.
        constexpr int L2BLK = 128;
        std::vector<int2> all;
        for (int c0 = 0; c0 < d->nComp; c0 += L2BLK)
            for (int t = 0; t < d->nTiles; ++t)
                for (int c = c0; c < std::min(c0 + L2BLK, d->nComp); ++c) {
                    if (!hit[c] || (sym && d->compTile[c] < t)) continue;
                    all.push_back(make_int2(c, t));
                }
        // this rank's share: a contiguous slice of the (equal-cost) block-pair list
        const size_t wb = all.size() * (size_t)shareRank / shareN, we = all.size() * (size_t)(shareRank + 1) / shareN;
        d->work.assign(all.begin() + wb, all.begin() + we);
        double fl = 0.0;
        for (const int2 &wk : d->work) {
            // useful FP64 flops of this CTA (padding excluded): main loop 2 nv no per (site, row),
            // one multiply per A element per column tile, epilogue (1 + 2 r_f) n_c per accumulator element
            const double ns = d->compNumSites[wk.x], nr = d->tileNumRows[wk.y];
            fl += ns * nr * (2.0 * d->nv * d->no + (double)d->nv * d->nOt + ns * d->no * (1.0 + 2.0 * d->rf));
        }
        d->emptyInRange.clear();
        if (shareRank == 0 && !d->expanded)
            for (int p : d->emptyP)
                if (p >= v.colBegin && p < v.colEnd) d->emptyInRange.push_back(p);
        if (!d->emptyInRange.empty()) {
            if (!d->d_emptyCols && cudaMalloc((void **)&d->d_emptyCols, std::max<size_t>(d->emptyP.size(), 1) * sizeof(int)) != cudaSuccess) {
                cudaGetLastError();
                set_error("cudaMalloc failed (empty-column list)");
                return QCDMET_E_ALLOC;
            }
            QC_CUDA(cudaMemcpyAsync(d->d_emptyCols, d->emptyInRange.data(), d->emptyInRange.size() * sizeof(int),
                                    cudaMemcpyHostToDevice, st));
        }
        if (d->work.size() > d->workCapacity) {
            cudaFree(d->d_work);
            d->d_work = nullptr;
            d->workCapacity = 0;
            if (cudaMalloc((void **)&d->d_work, d->work.size() * sizeof(int2)) != cudaSuccess) {
                cudaGetLastError();
                set_error("cudaMalloc of the DMMA work list failed");
                return QCDMET_E_ALLOC;
            }
            d->workCapacity = d->work.size();
        }
        if (!d->work.empty())
            QC_CUDA(cudaMemcpyAsync(d->d_work, d->work.data(), d->work.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
        d->launchBegin = v.colBegin;
        d->launchEnd = v.colEnd;
        d->launchSym = sym;
        d->launchFlops = fl;
        d->launchRank = shareRank;
        d->launchRanks = shareN;
    }
    double *target = v.J;
    if (d->expanded) {
        const size_t need = (size_t)v.Q * std::max(ce - cb, 1);
        if (need > d->JsubCapacity) {
            cudaFree(d->d_Jsub);
            d->d_Jsub = nullptr;
            d->JsubCapacity = 0;
            if (cudaMalloc((void **)&d->d_Jsub, need * sizeof(double)) != cudaSuccess) {
                cudaGetLastError();
                set_error("cudaMalloc of the sub-perturbation scratch (%zu doubles) failed", need);
                return QCDMET_E_ALLOC;
            }
            d->JsubCapacity = need;
        }
        target = d->d_Jsub;
    }
    ZeroArgs z;
    z.nDest = peers ? peers->nranks : 1;
    for (int r = 0; r < QCDMET_MAX_PEERS; ++r) z.J[r] = (peers && r < peers->nranks) ? peers->J[r] : (r == 0 ? target : nullptr);
    if (!d->emptyInRange.empty()) {
        const unsigned gx = (unsigned)std::min<long long>((v.Q + 255) / 256, 1024);
        zero_columns_kernel<<<dim3(gx, (unsigned)d->emptyInRange.size()), 256, 0, st>>>(z, d->d_emptyCols, v.colBegin, v.Q);
        QC_CUDA(cudaGetLastError());
        *launches += 1;
    }
    if (!d->work.empty()) {
        DmmaArgs a;
        a.nKc = d->nKc; a.nOt = d->nOt;
        a.colBegin = cb; a.colEnd = ce; a.sym = d->launchSym ? 1 : 0; a.Q = v.Q;
        a.jCap = (long long)v.Q * (ce - cb);
        a.Wt = d->d_Wt; a.PVt = d->d_PVt; a.CVt = d->d_CVt; a.POt = d->d_POt; a.COt = d->d_COt;
        a.work = d->d_work; a.nWork = (int)d->work.size();
        a.tileTab = d->d_tileTab; a.compTab = d->d_compTab;
        a.nP = d->nP; a.nTP = d->nTP; a.maxE = d->maxE;
        a.Mout = nullptr; a.mRows = 0; a.ldm = 0;
        a.nDest = z.nDest;
        for (int r = 0; r < QCDMET_MAX_PEERS; ++r) a.J[r] = z.J[r];
        // tuning aid: $QCDMET_B200_PROFILE=<file> runs the instrumented instantiation and appends the per-phase cycle
        // counts of the consumer warps (averages over CTAs) to <file>; never set in production or in the bench
        const char *profPath = getenv("QCDMET_B200_PROFILE");
        const unsigned gridForProf = (unsigned)std::min<size_t>(d->work.size(), (size_t)d->numSMs);
        a.prof = nullptr;
        if (profPath && d->rf == 4 && d->nj >= 6) {
            QC_CUDA(cudaMalloc((void **)&a.prof, (size_t)gridForProf * 64 * sizeof(long long)));
            QC_CUDA(cudaMemsetAsync(a.prof, 0, (size_t)gridForProf * 64 * sizeof(long long), st));
        }
        if (ev0) QC_CUDA(cudaEventRecord(ev0, st));
        // persistent CTAs, one per SM: CTA b executes items b, b + grid, ... (equal-cost items)
        const unsigned grid = (unsigned)std::min<size_t>(d->work.size(), (size_t)d->numSMs);
        // Tile-granular balancing (SPLIT kernel).  With W items on G SMs the busiest CTA executes ceil(W / G) items; when W
        // is not a multiple of G that is up to one item too many (N = 576: 10 for 9.24).  The SPLIT instantiation gives every
        // CTA a contiguous range of (item, column tile) units instead -- ceil(W nOt / G) tiles -- and an item cut by a range
        // boundary is completed by the CTA holding its first tiles, which adds the other CTA's per-tile contributions in
        // tile order: the same association as an unsplit item, so results stay bit-identical to the unsplit kernel and
        // independent of G.  Used when it removes at least 4 % of the busiest CTA's work (measured: 4.1 % of the kernel time at
        // N = 576 for an expected 6.7 %; nothing at N = 1000 for an expected 2.7 %) and the operand panels are small
        // enough to stay in L2 whatever the order (contiguous ranges give up the tile-sharing order of the strided list).
        bool split = false;
        {
            const long long W = (long long)d->work.size(), G = grid, nOt = d->nOt;
            const char *env = getenv("QCDMET_B200_SPLIT");
            const size_t footprint = ((size_t)d->nOt * d->nKc * d->wStage() + (size_t)d->nTiles * d->nKc * PV_STAGE +
                                      (size_t)d->nComp * d->nKc * CV_STAGE + (size_t)d->nTiles * d->nOt * d->posDoubles() +
                                      (size_t)d->nComp * d->nOt * d->cosDoubles()) * 8;
            if (G == d->numSMs && nOt >= 2 && nOt <= 8 && W % G != 0 && !(env && env[0] == '0')) {
                const long long before = ((W + G - 1) / G) * nOt, after = (W * nOt + G - 1) / G;
                split = (W * nOt) / G >= 2 * nOt && after * 100 <= before * 96 && footprint <= ((size_t)48 << 20);
            }
            if (env && env[0] == '1' && G == d->numSMs && nOt >= 2 && (W * nOt) / G >= 2 * nOt) split = true;   // forced (tests)
        }
        if (split) {
            const size_t nsDoubles = (size_t)2 * 128 * (RT * d->rf + 2);
            const size_t need = (size_t)grid * (d->nOt - 1) * nsDoubles;
            if (need > d->wsCapacity) {
                cudaFree(d->d_ws);
                d->d_ws = nullptr;
                d->wsCapacity = 0;
                if (cudaMalloc((void **)&d->d_ws, need * sizeof(double)) != cudaSuccess) {
                    cudaGetLastError();
                    split = false;                               // no room: the unsplit kernel needs no workspace
                } else d->wsCapacity = need;
            }
            if (split && !d->d_flags && cudaMalloc((void **)&d->d_flags, (size_t)d->numSMs * sizeof(unsigned int)) != cudaSuccess) {
                cudaGetLastError();
                d->d_flags = nullptr;
                split = false;
            }
            if (split) QC_CUDA(cudaMemsetAsync(d->d_flags, 0, (size_t)d->numSMs * sizeof(unsigned int), st));
        }
        d->launchSplit = split;
        a.ws = d->d_ws;
        a.flags = d->d_flags;
        int rc = QCDMET_E_STATE;
#define QC_LAUNCH_VARIANT(RFV, NJV)                                                                           \
        if (d->rf == RFV && d->nj == NJV) {                                                                   \
            using K = Cfg<RFV, NJV>;                                                                          \
            if (!d->attrSet) {                                                                                \
                QC_CUDA(cudaFuncSetAttribute(jac_dmma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             K::SMEM_BYTES));                                                 \
                d->attrSet = true;                                                                            \
            }                                                                                                 \
            if (profPath && RFV == 4 && NJV >= 6) {                                                           \
                QC_CUDA(cudaFuncSetAttribute(jac_dmma_kernel<Cfg<4, (NJV >= 6 ? NJV : 6)>, false, true>,      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES));    \
                jac_dmma_kernel<Cfg<4, (NJV >= 6 ? NJV : 6)>, false, true><<<grid, NTHREADS, K::SMEM_BYTES, st>>>(a); \
            } else if (split) {                                                                               \
                QC_CUDA(cudaFuncSetAttribute(jac_dmma_kernel<K, false, false, true>,                          \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES));    \
                jac_dmma_kernel<K, false, false, true><<<grid, NTHREADS, K::SMEM_BYTES, st>>>(a);             \
            } else                                                                                            \
                jac_dmma_kernel<K><<<grid, NTHREADS, K::SMEM_BYTES, st>>>(a);                                 \
            rc = QCDMET_OK;                                                                                   \
        }
        QC_LAUNCH_VARIANT(4, 8) QC_LAUNCH_VARIANT(4, 6) QC_LAUNCH_VARIANT(4, 4)
        QC_LAUNCH_VARIANT(8, 8) QC_LAUNCH_VARIANT(8, 6) QC_LAUNCH_VARIANT(8, 4)
#undef QC_LAUNCH_VARIANT
        QC_CHECK(rc == QCDMET_OK, QCDMET_E_STATE, "no DMMA kernel variant for rf=%d nj=%d", d->rf, d->nj);
        QC_CUDA(cudaGetLastError());
        if (ev1) QC_CUDA(cudaEventRecord(ev1, st));
        if (a.prof) {
            std::vector<long long> h((size_t)gridForProf * 64);
            QC_CUDA(cudaMemcpyAsync(h.data(), a.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
            QC_CUDA(cudaStreamSynchronize(st));
            cudaFree(a.prof);
            if (FILE *fp = fopen(profPath, "a")) {
                static const char *names[8] = {"wait_full", "stage_body", "wait_epi", "epilogue", "assembly", "-", "-", "loop_overhead"};
                fprintf(fp, "items %zu ctas %u nKc %d nOt %d nj %d\n", d->work.size(), gridForProf, d->nKc, d->nOt, d->nj);
                for (int grp = 0; grp < 2; ++grp) {
                    double sum[8] = {0};
                    for (unsigned b = 0; b < gridForProf; ++b)
                        for (int wq = 4 * grp; wq < 4 * grp + 4; ++wq)
                            for (int i = 0; i < 8; ++i) sum[i] += (double)h[((size_t)b * 8 + wq) * 8 + i];
                    double tot = 0;
                    for (int i = 0; i < 8; ++i) tot += sum[i];
                    fprintf(fp, "  warps %d-%d: total %.0f cycles/warp;", 4 * grp, 4 * grp + 3, tot / (4.0 * gridForProf));
                    for (int i = 0; i < 8; ++i)
                        if (names[i][0] != '-') fprintf(fp, " %s %.1f%%", names[i], 100.0 * sum[i] / tot);
                    fprintf(fp, "\n");
                }
                fclose(fp);
            }
        }
        *launches += 1;
        *flops = d->launchFlops;
    }
    if (d->expanded) {
        // every sub-column in [cb, ce) has been written by exactly one CTA (each sub-perturbation is non-empty)
        const unsigned gx = (unsigned)std::min<long long>((v.Q + 255) / 256, 1024);
        sum_subcolumns_kernel<<<dim3(gx, (unsigned)ncols), 256, 0, st>>>(d->d_Jsub, d->d_subStart, v.colBegin, cb, v.Q, v.J);
        QC_CUDA(cudaGetLastError());
        *launches += 1;
    }
    return QCDMET_OK;
}

bool dmma_last_launch_symmetric(const DmmaPlanData *d) { return d && d->launchSym; }
bool dmma_last_launch_split(const DmmaPlanData *d) { return d && d->launchSplit; }
bool dmma_plan_expanded(const DmmaPlanData *d) { return d && d->expanded; }

// Host-only description of the schedule the DMMA kernel would use (no device needed): tests and tooling.
void dmma_schedule_info(int N, int no, int P, int F, const int *h1start, const int *h1row, const int *h1col,
                        const int *rowStart, const int *rowIdx, int *info, std::string *why)
{
    Schedule sc = build_schedule(N, no, P, F, h1start, h1row, h1col, rowStart, rowIdx);
    const DmmaPlanData &d = sc.d;
    info[0] = sc.ok ? 1 : 0;
    info[1] = d.rf; info[2] = d.nj; info[3] = d.nComp; info[4] = d.nTiles;
    info[5] = d.symOk ? 1 : 0; info[6] = d.expanded ? 1 : 0; info[7] = d.P2;
    // block pairs the full-range launch would execute (symmetric mode halves them)
    long long pairs = 0;
    if (sc.ok)
        for (int t = 0; t < d.nTiles; ++t)
            for (int c = 0; c < d.nComp; ++c)
                if (!(d.symOk && d.compTile[c] < t)) ++pairs;
    info[8] = (int)std::min<long long>(pairs, 2147483647LL);
    info[9] = d.nOt;
    if (why) *why = sc.why;
}

// ------------------------------------------------------------------------------------------
// Large fragments / large site groups (polyyne- and sn2-like shapes, the compat full response):
// the same DMMA main loop produces the panels M_r[a, o] for every site r that occurs as an element
// row and every fragment row a, stored to HBM (panel mode above); kernels_contract.cu then contracts,
// per perturbation, Y_p[a,o] = sum_{(r,c) in p} M_r[a,o] Co[c,o] with the occupied rows of the same
// fragment.  2 |sites| S nv no + P no (2 nnz_p S + 4 Q) flop instead of P (2 S nv no + ...).
// ------------------------------------------------------------------------------------------
struct FactPlanData {
    DmmaPlanData sched;                 // components = chunks of 4 used sites, tiles = chunks of 32 rows
    std::vector<int> siteToU;           // [N] index of the site in the panel array, -1 if it has no element row
    std::vector<int> elU;               // [nnz] panel index of the row site of every element
    int *d_elU = nullptr;
    double *d_M = nullptr;              // [4 nComp][S][ldo]
    size_t mDoubles = 0;
    int numSites = 0;
    int maxNnz = 0;                     // longest perturbation of the plan
    ContractPlan *contract = nullptr;
};

namespace {
int fact_pick_nj(int no)
{
    int best = 8;
    long long bestPad = -1;
    for (int nj : {8, 6, 4}) {
        const long long bn = 16 * nj, pad = (no + bn - 1) / bn * bn;
        if (bestPad < 0 || pad < bestPad) { bestPad = pad; best = nj; }
    }
    return best;
}
int fact_count_sites(int N, int P, const int *h1start, const int *h1row, std::vector<int> *siteToU)
{
    std::vector<char> used(N, 0);
    for (int k = 0; k < h1start[P]; ++k) used[h1row[k]] = 1;
    int n = 0;
    if (siteToU) siteToU->assign(N, -1);
    for (int i = 0; i < N; ++i)
        if (used[i]) { if (siteToU) (*siteToU)[i] = n; ++n; }
    return n;
}
}  // namespace

bool fact_eligible(int P, const int *h1start, std::string *why)
{
    for (int p = 0; p < P; ++p)
        if (h1start[p + 1] - h1start[p] > QCDMET_CONTRACT_MAX_NNZ) {
            if (why) *why = "a perturbation has more than 64 elements";
            return false;
        }
    return true;
}

// device memory the factorised path needs on top of the plan's own buffers
size_t fact_plan_bytes(int N, int no, int P, int S, const int *h1start, const int *h1row)
{
    const int sites = fact_count_sites(N, P, h1start, h1row, nullptr);
    const size_t nComp = (sites + RT - 1) / RT, nTiles = (S + AT - 1) / AT, nv = N - no;
    const size_t nKc = (nv + KC - 1) / KC, bn = 16 * fact_pick_nj(no), nOt = (no + bn - 1) / bn;
    const size_t ldo = round_up(no, 16);
    return (nComp * RT * (size_t)S * ldo + nOt * nKc * KC * bn + nTiles * nKc * PV_STAGE + nComp * nKc * CV_STAGE) * 8;
}

int fact_plan_build(FactPlanData **out, int N, int no, int P, int F, const int *rowStart, int ldo,
                    const int *h1start, const int *h1row)
{
    *out = nullptr;
    const int S = F > 0 ? rowStart[F] : 0;
    FactPlanData *f = new FactPlanData();
    DmmaPlanData &d = f->sched;
    d.N = N; d.no = no; d.nv = N - no; d.P = P; d.S = S;
    f->numSites = fact_count_sites(N, P, h1start, h1row, &f->siteToU);
    d.nComp = (f->numSites + RT - 1) / RT;
    d.compSites.assign((size_t)d.nComp * RT, -1);
    for (int i = 0; i < N; ++i)
        if (f->siteToU[i] >= 0) d.compSites[f->siteToU[i]] = i;
    for (int p = 0; p < P; ++p) f->maxNnz = std::max(f->maxNnz, h1start[p + 1] - h1start[p]);
    f->elU.resize(h1start[P]);
    for (int k = 0; k < h1start[P]; ++k) f->elU[k] = f->siteToU[h1row[k]];
    d.compNumSites.assign(d.nComp, RT);
    if (f->numSites % RT) d.compNumSites.back() = f->numSites % RT;
    d.nTiles = (S + AT - 1) / AT;
    d.tileRow.assign((size_t)d.nTiles * AT, -1);
    for (int a = 0; a < S; ++a) d.tileRow[a] = a;
    d.tileNumRows.assign(d.nTiles, AT);
    if (S % AT) d.tileNumRows.back() = S % AT;
    // tables the panel mode loads but never uses (kept at their legal size for Cfg<4, .>)
    d.rf = 4;
    d.compP.assign((size_t)d.nComp * MAXP, -1);
    d.compEnt.assign((size_t)d.nComp * MAXP * MAXE, 0xFF);
    d.tilePartner.assign((size_t)d.nTiles * AT * 4, (unsigned char)AT);
    d.tileSelf.assign((size_t)d.nTiles * AT, 0);
    d.tileQ.assign((size_t)d.nTiles * AT, 0);
    d.nj = fact_pick_nj(no);
    d.nKc = (d.nv + KC - 1) / KC;
    d.nOt = (d.no + d.bn() - 1) / d.bn();
    f->mDoubles = (size_t)d.nComp * RT * S * ldo;
    int rc = QCDMET_OK;
    pack_item_tables(d);
    if ((rc = query_sms(d)) || (rc = upload(&d.d_compSites, d.compSites)) || (rc = upload(&d.d_tileRow, d.tileRow)) ||
        (rc = upload(&d.d_tileTab, d.tileTab)) || (rc = upload(&d.d_compTab, d.compTab)) ||
        (rc = upload(&f->d_elU, f->elU)) ||
        (rc = contract_plan_build(&f->contract, F, rowStart))) {
        fact_plan_free(f);
        return rc;
    }
    d.work.clear();
    for (int t = 0; t < d.nTiles; ++t)
        for (int c = 0; c < d.nComp; ++c) d.work.push_back(make_int2(c, t));
    const size_t one = 1;
    if (cudaMalloc((void **)&d.d_work, std::max(d.work.size(), one) * sizeof(int2)) != cudaSuccess ||
        cudaMalloc((void **)&d.d_Wt, std::max((size_t)d.nOt * d.nKc * d.wStage(), one) * 8) != cudaSuccess ||
        cudaMalloc((void **)&d.d_PVt, std::max((size_t)d.nTiles * d.nKc * PV_STAGE, one) * 8) != cudaSuccess ||
        cudaMalloc((void **)&d.d_CVt, std::max((size_t)d.nComp * d.nKc * CV_STAGE, one) * 8) != cudaSuccess ||
        cudaMalloc((void **)&f->d_M, std::max(f->mDoubles, one) * 8) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of the factorised-path buffers failed (%zu bytes for the panels M)", f->mDoubles * 8);
        fact_plan_free(f);
        return QCDMET_E_ALLOC;
    }
    if (!d.work.empty())
        QC_CUDA(cudaMemcpy(d.d_work, d.work.data(), d.work.size() * sizeof(int2), cudaMemcpyHostToDevice));
    *out = f;
    return QCDMET_OK;
}

void fact_plan_free(FactPlanData *f)
{
    if (!f) return;
    DmmaPlanData &d = f->sched;
    cudaFree(d.d_compSites); cudaFree(d.d_tileRow); cudaFree(d.d_tileTab); cudaFree(d.d_compTab); cudaFree(d.d_work);
    cudaFree(d.d_Wt); cudaFree(d.d_PVt); cudaFree(d.d_CVt);
    cudaFree(f->d_elU); cudaFree(f->d_M);
    cudaGetLastError();
    contract_plan_free(f->contract);
    delete f;
}

// tiled operands of the panel kernel from the current orbitals
int fact_prepare_orbitals(FactPlanData *f, const JacView &v, cudaStream_t st)
{
    DmmaPlanData *d = &f->sched;
    const long long wt = (long long)d->nOt * d->nKc * d->wStage(), pvt = (long long)d->nTiles * d->nKc * PV_STAGE,
                    cvt = (long long)d->nComp * d->nKc * CV_STAGE;
    if (wt > 0) {
        build_wt_kernel<<<(unsigned)((wt + 255) / 256), 256, 0, st>>>(v.e, v.no, v.nv, d->nKc, d->bn(), wt, d->d_Wt);
        QC_CUDA(cudaGetLastError());
    }
    if (pvt > 0) {
        build_pvt_kernel<<<(unsigned)((pvt + 255) / 256), 256, 0, st>>>(v.PV, v.ldv, v.nv, d->nKc, d->d_tileRow, pvt, d->d_PVt);
        QC_CUDA(cudaGetLastError());
    }
    if (cvt > 0) {
        build_cvt_kernel<<<(unsigned)((cvt + 255) / 256), 256, 0, st>>>(v.CvT, v.ldv, v.nv, d->nKc, d->d_compSites, cvt, d->d_CVt);
        QC_CUDA(cudaGetLastError());
    }
    return QCDMET_OK;
}

// Panels M for all element-row sites (DMMA main loop, panel mode), then the per-perturbation contraction
// for the columns of the view.  ev0/ev1 bracket both kernels.  The panels are recomputed by every call
// (2 |sites| S nv no flop: a column-sharded rank repeats them, they are never served from a cache).
int launch_jacobian_factorised(FactPlanData *f, const JacView &v, long long nnzInRange, cudaStream_t st,
                               int *launches, double *flops, cudaEvent_t ev0, cudaEvent_t ev1)
{
    DmmaPlanData *d = &f->sched;
    *launches = 0;
    *flops = 0.0;
    if (v.colEnd <= v.colBegin || v.S == 0) return QCDMET_OK;
    if (ev0) QC_CUDA(cudaEventRecord(ev0, st));
    if (!d->work.empty()) {
        DmmaArgs a;
        memset(&a, 0, sizeof(a));
        a.nKc = d->nKc; a.nOt = d->nOt;
        a.Wt = d->d_Wt; a.PVt = d->d_PVt; a.CVt = d->d_CVt;
        a.work = d->d_work; a.nWork = (int)d->work.size();
        a.tileTab = d->d_tileTab; a.compTab = d->d_compTab;
        a.Mout = f->d_M; a.mRows = v.S; a.ldm = v.ldo;
        a.prof = nullptr;
        a.ws = nullptr; a.flags = nullptr;
        a.nDest = 0;
        const unsigned grid = (unsigned)std::min<size_t>(d->work.size(), (size_t)d->numSMs);
        int rc = QCDMET_E_STATE;
        // per device (function attributes live in the context of the current device): set on every launch
#define QC_LAUNCH_PANEL(NJV)                                                                                  \
        if (d->nj == NJV) {                                                                                   \
            using K = Cfg<4, NJV>;                                                                            \
            QC_CUDA(cudaFuncSetAttribute(jac_dmma_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         K::SMEM_BYTES));                                                     \
            jac_dmma_kernel<K, true><<<grid, NTHREADS, K::SMEM_BYTES, st>>>(a);                               \
            rc = QCDMET_OK;                                                                                   \
        }
        QC_LAUNCH_PANEL(8) QC_LAUNCH_PANEL(6) QC_LAUNCH_PANEL(4)
#undef QC_LAUNCH_PANEL
        QC_CHECK(rc == QCDMET_OK, QCDMET_E_STATE, "no panel kernel variant for nj=%d", d->nj);
        QC_CUDA(cudaGetLastError());
        *launches += 1;
        // useful flops: main loop 2 nv no per (site, row) + one multiply per A element per column tile
        *flops += (double)f->numSites * (double)v.S * (2.0 * d->nv * d->no + (double)d->nv * d->nOt);
    }
    int l2 = 0;
    double f2 = 0.0;
    QC_TRY(launch_contract_panels(f->contract, v, f->d_M, f->d_elU, nnzInRange, f->maxNnz <= 2, st, &l2, &f2));
    if (ev1) QC_CUDA(cudaEventRecord(ev1, st));
    *launches += l2;
    *flops += f2;
    return QCDMET_OK;
}

}  // namespace qcdmet
