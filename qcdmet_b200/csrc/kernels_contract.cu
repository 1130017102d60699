// kernels_contract.cu -- per-perturbation contraction of the site panels M_r (factorised path for
// fragments of more than 8 rows: polyyne- / sn2-like fits and the compat full response).
//
// kernels_dmma.cu (panel mode) leaves  M_r[a,o] = sum_v PV[a,v] Cv[r,v] W[v,o]  in HBM for every site r
// that occurs as an element row.  Per column p (lib/rhf_response.cc:82-108 restricted to fragment blocks):
//   Y_p[a,o]      = sum_{(r,c) in H1_p} M_r[a,o] Co[c,o]
//   J[(f;a,b),p]  = 2 sum_o ( Y_p[a,o] PO[b,o] + PO[a,o] Y_p[b,o] )        a, b rows of fragment f
// a symmetric rank-2k update per fragment, on the FP64 tensor pipe (DMMA.8x8x4).  Work decomposition,
// built once per plan on the host:
//   block = 8 consecutive rows of one fragment (the last block of a fragment may be ragged)
//   strip = (block b1; up to 4 consecutive blocks b2 >= b1 of the same fragment): the A fragments of b1 are
//           loaded once per k-step and serve all pairs of the strip; 2 DMMA per pair and k-step
//   tile  = <= 16 blocks (128 row slots) and <= 24 strips, at most 3 per warp: whole fragments packed
//           together, or a pair of 8-block slices of a large fragment.  One CTA = (tile, 8 consecutive columns).
// Every (a,b) of every block has exactly one owner and J[a,b], J[b,a] are stored from the same register:
// no atomics, blocks exactly symmetric, run-to-run bit-reproducible.
#include "common.cuh"
#include <algorithm>
#include <cstdint>

namespace qcdmet {

namespace {
constexpr int CB8 = 8;                 // rows per block
constexpr int CMAXB = 16;              // blocks per tile
constexpr int CSLOTS = CB8 * CMAXB;    // row slots per tile
constexpr int CNT = 256;               // threads per CTA
constexpr int CWARPS = CNT / 32;
constexpr int CMAXS = QCDMET_CONTRACT_STRIPS_PER_WARP;
constexpr int CMAXN2 = 4;              // block pairs per strip
constexpr int CTO = 32;                // occupied orbitals per staged chunk
// Shared-memory layout of a staged chunk: Ys[slot][o], row stride CLD = 4 (mod 16) doubles: the 8 rows x 4
// orbitals of a DMMA operand fragment (LDS.64, half-warp = 4 rows x 4 orbitals) and the staging stores
// (STS.128, quarter-warp = 4 rows x 2 orbital pairs) both touch every bank once.
constexpr int CLD = CTO + 4;           // 36
constexpr int CBUF = CSLOTS * CLD;     // doubles per staged array
constexpr int CCB = 8;                 // columns per CTA
constexpr int CSTG = 4;                // staging tasks in flight per thread (loads issued before they are used)
constexpr int CMAXE = QCDMET_CONTRACT_MAX_NNZ;
constexpr int RES_MAXCOLS = 128;       // columns per CTA of the resident variant (descriptor table in shared memory)
constexpr int RMAXB = 8;               // blocks per tile of the 64-slot resident variant (two CTAs per SM)
constexpr int CSMEM = 2 * CBUF * 8 + CCB * CMAXE * 16 + CCB * 4 + CSLOTS * 4;
// warp-specialised variant (contract_panels_ws_kernel): 4 staging warps + 8 DMMA warps, 2 stages
constexpr int WS_STAGE_WARPS = 4;
constexpr int WS_NT = 32 * (WS_STAGE_WARPS + CWARPS);       // 384
constexpr int WS_TASKS = CSLOTS / 4 * 2 / WS_STAGE_WARPS;   // staging tasks per thread and chunk: 16
constexpr int WS_BATCH = 8;                                 // of which in flight (the staging warps hold no accumulators)
constexpr int WS_SMEM = 4 * CBUF * 8 + CCB * CMAXE * 16 + CCB * 4 + CSLOTS * 4 + 4 * 8;
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
        "CWAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CWAIT_DONE;\n"
        "bra CWAIT_LOOP;\n"
        "CWAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
}  // namespace

// ------------------------------------------------------------------------------------------
// host: schedule
// ------------------------------------------------------------------------------------------
void contract_schedule_build(int F, const int *rowStart, ContractSchedule *out, int maxBlocks)
{
    if (maxBlocks <= 0) maxBlocks = CMAXB;
    const int slice = maxBlocks / 2;
    ContractSchedule &s = *out;
    s.tileInfo.clear(); s.slotRow.clear(); s.strips.clear();
    s.nTiles = 0; s.validSlots = 0; s.blockPairs = 0;
    struct Strip { int b1, b2, n2; };
    std::vector<Strip> cur;            // strips of the tile under construction (tile-local block indices)
    int blocks = 0, slotBegin = 0;
    auto flush = [&]() {
        if (blocks > 0) {
            // longest strips first, each to the least loaded warp that still has a free slot
            std::vector<int> order(cur.size());
            for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cur[a].n2 > cur[b].n2; });
            int load[CWARPS] = {0}, cnt[CWARPS] = {0};
            std::vector<int> tab(CWARPS * CMAXS, -1);
            for (int i : order) {
                int best = -1;
                for (int w = 0; w < CWARPS; ++w)
                    if (cnt[w] < CMAXS && (best < 0 || load[w] < load[best])) best = w;
                tab[best * CMAXS + cnt[best]++] = cur[i].b1 | (cur[i].b2 << 8) | (cur[i].n2 << 16);
                load[best] += cur[i].n2;
                s.blockPairs += cur[i].n2;
            }
            s.tileInfo.push_back(slotBegin); s.tileInfo.push_back(blocks * CB8);
            s.tileInfo.push_back((int)s.strips.size()); s.tileInfo.push_back((int)cur.size());
            s.strips.insert(s.strips.end(), tab.begin(), tab.end());
            ++s.nTiles;
        }
        cur.clear();
        blocks = 0;
        slotBegin = (int)s.slotRow.size();
    };
    auto push_blocks = [&](int rs, int rf, int b0, int nb) {
        for (int b = b0; b < b0 + nb; ++b)
            for (int i = 0; i < CB8; ++i) {
                const int al = b * CB8 + i;
                s.slotRow.push_back(al < rf ? rs + al : -1);
                if (al < rf) ++s.validSlots;
            }
    };
    // strips covering the pairs (a0 + x, c0 + y), x < na, y < nc (y >= x only when diag)
    auto add_strips = [&](int a0, int na, int c0, int nc, bool diag, std::vector<Strip> *dst) {
        for (int x = 0; x < na; ++x)
            for (int y = diag ? x : 0; y < nc; y += CMAXN2)
                dst->push_back({a0 + x, c0 + y, std::min(CMAXN2, nc - y)});
    };
    for (int f = 0; f < F; ++f) {
        const int rs = rowStart[f], rf = rowStart[f + 1] - rs;
        if (rf <= 0) continue;
        const int nb = (rf + CB8 - 1) / CB8;
        std::vector<Strip> mine;
        if (nb <= maxBlocks) add_strips(0, nb, 0, nb, true, &mine);
        if (nb <= maxBlocks && (int)mine.size() <= CWARPS * CMAXS) {
            // whole fragment in one tile, packed with its neighbours
            if (blocks + nb > maxBlocks || cur.size() + mine.size() > (size_t)(CWARPS * CMAXS)) flush();
            for (const Strip &st : mine) cur.push_back({st.b1 + blocks, st.b2 + blocks, st.n2});
            push_blocks(rs, rf, 0, nb);
            blocks += nb;
        } else {
            // large fragment: one tile per pair of slices of maxBlocks / 2 blocks (si <= sj)
            flush();
            const int ns = (nb + slice - 1) / slice;
            for (int si = 0; si < ns; ++si)
                for (int sj = si; sj < ns; ++sj) {
                    const int bi0 = si * slice, bin = std::min(slice, nb - bi0);
                    const int bj0 = sj * slice, bjn = std::min(slice, nb - bj0);
                    push_blocks(rs, rf, bi0, bin);
                    if (sj == si) {
                        add_strips(0, bin, 0, bin, true, &cur);
                        blocks = bin;
                    } else {
                        push_blocks(rs, rf, bj0, bjn);
                        add_strips(0, bin, bin, bjn, false, &cur);
                        blocks = bin + bjn;
                    }
                    flush();
                }
        }
    }
    flush();
}

// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------
namespace {
// One strip over the k-steps of a staged chunk: block b1 against N2 consecutive blocks b2.
// D[b1,b2] += Y[b1] PO[b2]^T + PO[b1] Y[b2]^T; the two products of a pair are issued N2 DMMAs apart.
template <int N2, int LD = CLD, bool SPLIT = false, int UNROLL = 2>
__device__ __forceinline__ void strip_steps(double (&acc)[CMAXN2][2], const double *__restrict__ ya,
                                            const double *__restrict__ pa, const double *__restrict__ yb,
                                            const double *__restrict__ pb, int nk)
{
    // SPLIT: the two products of a pair accumulate in separate registers (two independent DMMA chains per pair
    // instead of one: a strip of one pair no longer issues dependent DMMAs back to back) and are added at the end
    double c2[SPLIT ? N2 : 1][2];
    if constexpr (SPLIT) {
#pragma unroll
        for (int j = 0; j < N2; ++j) c2[j][0] = c2[j][1] = 0.0;
    }
#pragma unroll UNROLL
    for (int kk = 0; kk < nk; ++kk) {
        const double aY = ya[4 * kk], aP = pa[4 * kk];
        double bP[N2], bY[N2];
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            bP[j] = pb[j * CB8 * LD + 4 * kk];
            bY[j] = yb[j * CB8 * LD + 4 * kk];
        }
#pragma unroll
        for (int j = 0; j < N2; ++j) dmma884(acc[j][0], acc[j][1], aY, bP[j]);
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            if constexpr (SPLIT) dmma884(c2[j][0], c2[j][1], aP, bY[j]);
            else dmma884(acc[j][0], acc[j][1], aP, bY[j]);
        }
    }
    if constexpr (SPLIT) {
#pragma unroll
        for (int j = 0; j < N2; ++j) { acc[j][0] += c2[j][0]; acc[j][1] += c2[j][1]; }
    }
}
}  // namespace

// NNZ2: every perturbation of the plan has at most 2 elements (all DMET lists of src/dmet.py:118-159 outside
// the translation-invariant case): both panel loads of a task are prefetched, no element loop.
template <bool NNZ2>
__global__ void __launch_bounds__(CNT, 2)
contract_panels_kernel(const JacView v, const double *__restrict__ M, const int *__restrict__ elU,
                       const int4 *__restrict__ tileInfo, const int *__restrict__ slotRowG,
                       const int *__restrict__ stripsG)
{
    extern __shared__ __align__(16) unsigned char csm[];
    double *Ys = reinterpret_cast<double *>(csm);                    // Ys[slot][o]  Y_p of the tile's rows
    double *Ps = Ys + CBUF;                                          // Ps[slot][o]  PO of the same rows
    long long *elM = reinterpret_cast<long long *>(Ps + CBUF);       // [CCB][CMAXE] offset of the panel M_r of element e
    long long *elC = elM + CCB * CMAXE;                              // [CCB][CMAXE] offset of the row Co[c,:] of element e
    int *nnzs = reinterpret_cast<int *>(elC + CCB * CMAXE);          // [CCB] elements of the column
    int *srow = nnzs + CCB;                                          // [CSLOTS] fragment row of the slot or -1

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int4 ti = tileInfo[blockIdx.x];            // slotBegin, nSlots, first strip-table entry, nStrips
    const int pl0 = blockIdx.y * CCB;                // first local column of the batch
    const int ncb = min(CCB, v.colEnd - v.colBegin - pl0);

    if (tid < CSLOTS) srow[tid] = tid < ti.y ? slotRowG[ti.x + tid] : -1;
    for (int i = tid; i < ncb * CMAXE; i += CNT) {
        const int cb = i / CMAXE, e = i % CMAXE;
        const int p = v.colBegin + pl0 + cb;
        const int eb = v.h1start[p], n = v.h1start[p + 1] - eb;   // n <= CMAXE (checked at plan creation)
        if (e == 0) nnzs[cb] = n;
        if (e < n) {
            elM[i] = (long long)elU[eb + e] * v.S * v.ldo;        // elU[k] = panel index of the site h1row[k]
            elC[i] = (long long)v.h1col[eb + e] * v.ldo;
        }
    }
    // padding slots of ragged blocks stay zero for the whole CTA: staging never touches them
    for (int i = tid; i < CBUF; i += CNT) *reinterpret_cast<double2 *>(Ys + 2 * i) = make_double2(0.0, 0.0);
    // this warp's strips: b1 | b2 << 8 | n2 << 16, -1 = none
    int sw[CMAXS];
#pragma unroll
    for (int s = 0; s < CMAXS; ++s) sw[s] = stripsG[ti.z + warp * CMAXS + s];
    double acc[CMAXS][CMAXN2][2];
#pragma unroll
    for (int s = 0; s < CMAXS; ++s)
#pragma unroll
        for (int j = 0; j < CMAXN2; ++j) acc[s][j][0] = acc[s][j][1] = 0.0;
    __syncthreads();

    const int g4 = lane >> 2, t4 = lane & 3;         // DMMA fragment coordinates: row / column, k index
    constexpr int STG = NNZ2 ? CSTG : 2;             // staging tasks in flight per thread (register budget)
    static_assert(8 % STG == 0, "the 8 tasks of a thread are staged in equal batches");
    // Staging geometry, fixed for the CTA.  A task = (4 consecutive slots, 16-orbital half of the chunk);
    // lane = (slot of the group, orbital pair): a warp load is 4 rows x 128 contiguous bytes, a warp store
    // (STS.128) is conflict-free.  Warp w serves half w & 1 of the groups (w >> 1) + 4 i, i < 8.
    const int ls = lane & 3, lo = lane >> 2;
    const int half = warp & 1;
    const int colo = 16 * half + 2 * lo;             // orbital of this lane inside the chunk
    const int dst0 = (4 * (warp >> 1) + ls) * CLD + colo;
    int roff[8];                                     // row * ldo + colo of task i; < 0: padding slot or no task
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int slot = 4 * ((warp >> 1) + 4 * i) + ls;
        const int row = slot < ti.y ? srow[slot] : -1;
        roff[i] = row >= 0 ? row * v.ldo + colo : -1;            // S ldo < 2^31 (checked at plan creation)
    }
    int offA[CMAXS], offB[CMAXS];                    // operand fragments of the strips in a staged array
#pragma unroll
    for (int s = 0; s < CMAXS; ++s) {
        offA[s] = ((sw[s] & 0xFF) * CB8 + g4) * CLD + t4;
        offB[s] = (((sw[s] >> 8) & 0xFF) * CB8 + g4) * CLD + t4;
    }

    for (int cb = 0; cb < ncb; ++cb) {
        const int nnz = nnzs[cb];
        const long long *eM = elM + cb * CMAXE, *eC = elC + cb * CMAXE;
        const double *pM0 = M + (nnz > 0 ? eM[0] : 0), *pM1 = M + (nnz > 1 ? eM[1] : 0);
        const double *pC0 = v.CoT + (nnz > 0 ? eC[0] : 0), *pC1 = v.CoT + (nnz > 1 ? eC[1] : 0);
        for (int o0 = 0; o0 < v.no; o0 += CTO) {
            if (o0 + 16 * half < v.ldo) {            // this warp's half exists (ldo is a multiple of 16; panels zero padded)
                double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
                if (nnz > 0) c0 = ld2(pC0 + o0 + colo);
                if (nnz > 1) c1 = ld2(pC1 + o0 + colo);
#pragma unroll
                for (int i0 = 0; i0 < 8; i0 += STG) {
                    double2 m0[STG], m1[STG], po[STG];
#pragma unroll
                    for (int k = 0; k < STG; ++k) {
                        m0[k] = m1[k] = po[k] = make_double2(0.0, 0.0);
                        if (roff[i0 + k] >= 0) {
                            po[k] = ld2(v.PO + roff[i0 + k] + o0);
                            if (nnz > 0) m0[k] = ld2(pM0 + roff[i0 + k] + o0);
                            if (nnz > 1) m1[k] = ld2(pM1 + roff[i0 + k] + o0);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < STG; ++k) {
                        if (roff[i0 + k] < 0) continue;
                        double2 y;
                        y.x = fma(m1[k].x, c1.x, m0[k].x * c0.x);
                        y.y = fma(m1[k].y, c1.y, m0[k].y * c0.y);
                        if constexpr (!NNZ2) {
                            for (int e = 2; e < nnz; ++e) {      // longer lists (translation-invariant, arbitrary CSR)
                                const double2 m = ld2(M + eM[e] + roff[i0 + k] + o0), c = ld2(v.CoT + eC[e] + o0 + colo);
                                y.x = fma(m.x, c.x, y.x); y.y = fma(m.y, c.y, y.y);
                            }
                        }
                        *reinterpret_cast<double2 *>(Ys + dst0 + (i0 + k) * 16 * CLD) = y;
                        *reinterpret_cast<double2 *>(Ps + dst0 + (i0 + k) * 16 * CLD) = po[k];
                    }
                }
            }
            __syncthreads();
            // k-steps of 4 orbitals; the orbitals of the last step beyond no are zero in the panels
            const int nk = (min(CTO, v.no - o0) + 3) / 4;
#pragma unroll
            for (int s = 0; s < CMAXS; ++s) {
                if (sw[s] < 0) continue;                 // warp-uniform
                const double *ya = Ys + offA[s], *pa = Ps + offA[s], *yb = Ys + offB[s], *pb = Ps + offB[s];
                switch (sw[s] >> 16) {                   // warp-uniform
                case 4: strip_steps<4>(acc[s], ya, pa, yb, pb, nk); break;
                case 3: strip_steps<3>(acc[s], ya, pa, yb, pb, nk); break;
                case 2: strip_steps<2>(acc[s], ya, pa, yb, pb, nk); break;
                default: strip_steps<1>(acc[s], ya, pa, yb, pb, nk); break;
                }
            }
            __syncthreads();
        }
        // column done: J[a,b] and J[b,a] from the same register, then restart the sums.  Accumulator layout of
        // DMMA.8x8x4: the lane holds D[g4][2 t4] and D[g4][2 t4 + 1].
        double *Jc = v.J + (size_t)v.Q * (pl0 + cb);
#pragma unroll
        for (int s = 0; s < CMAXS; ++s) {
            if (sw[s] < 0) continue;
            const int b1 = sw[s] & 0xFF, b2 = (sw[s] >> 8) & 0xFF, n2 = sw[s] >> 16;
            // fragment of the strip's rows (the first row of a block always exists)
            const int f = v.rowFrag[srow[b1 * CB8]];
            const int rs = v.rowStart[f], rf = v.rowStart[f + 1] - rs;
            const int ra = srow[b1 * CB8 + g4];
            double *Jp = Jc + v.qoff[f];
#pragma unroll
            for (int j = 0; j < CMAXN2; ++j) {
                if (j < n2) {
                    const bool diag = (b2 + j == b1);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int rb = srow[(b2 + j) * CB8 + 2 * t4 + e];
                        if (ra >= 0 && rb >= 0 && !(diag && 2 * t4 + e < g4)) {
                            const long long al = ra - rs, bl = rb - rs;
                            const double val = 2.0 * acc[s][j][e];
                            QC_DEV_ASSERT(al >= 0 && bl >= 0 && al < rf && bl < rf);
                            Jp[al + rf * bl] = val;
                            Jp[bl + rf * al] = val;
                        }
                    }
                }
                acc[s][j][0] = acc[s][j][1] = 0.0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Resident variant for small occupied spaces (ldo <= 64: the polyyne- and sn2-like fits) and DMET lists
// (<= 2 elements per perturbation).  contract_panels_kernel re-stages three panel chunks (M of both elements, PO)
// from L2 for every column; here a CTA owns a tile for a long run of consecutive columns and keeps in shared
// memory, over the WHOLE occupied range: the PO rows of the tile (loaded once per CTA) and the panel of the
// column's first element (consecutive columns (r,r), (r,r+1), ... share the row site r: loaded once per run).
// Only the second element's panel is streamed -- prefetched into registers while the DMMA strips of the previous
// column run -- so the L2 -> SM traffic per column drops from 3 panel tiles to 1 + 1/(run length) + 1/(columns per
// CTA).  CTAs are sized by the host so that every
// CTA has the same number of block-pair DMMAs.  Tiles of 64 row slots (8 blocks): 3 x 64 x (ldo + 4) doubles, so two
// CTAs share an SM and one runs its DMMA strips while the other stages Y or stores.
// ------------------------------------------------------------------------------------------
template <int LDO, int RSLOTS>
__global__ void __launch_bounds__(CNT, RSLOTS <= 64 ? 2 : 1)
contract_resident_kernel(const JacView v, const double *__restrict__ M, const int *__restrict__ elU,
                         const int4 *__restrict__ tileInfo, const int *__restrict__ slotRowG,
                         const int *__restrict__ stripsG, const int4 *__restrict__ ctaTab)
{
    constexpr int RLD = LDO + 4;                   // row stride: 4 (mod 16) doubles, as CLD
    constexpr int RBUF = RSLOTS * RLD;
    constexpr int QP = LDO / 2;                    // orbital pairs per row
    constexpr int NIT = RSLOTS * QP / CNT;         // (slot, orbital pair) items per thread: 4 / 8 / 12 / 16
    static_assert(RSLOTS * QP % CNT == 0, "items divide evenly");
    extern __shared__ __align__(16) unsigned char csm[];
    double *Ys = reinterpret_cast<double *>(csm);          // Y_p of the tile's rows, whole occupied range
    double *Ps = Ys + RBUF;                                // PO of the same rows
    double *Mr = Ps + RBUF;                                // panel of the first element's row site
    double *coS = Mr + RBUF;                               // [2 buffers][2 elements][LDO] rows of C_o that scale the panels
    long long *dPan = reinterpret_cast<long long *>(coS + 4 * LDO);   // [RES_MAXCOLS][2] panel offsets of the elements, -1 = none
    int *dCol = reinterpret_cast<int *>(dPan + 2 * RES_MAXCOLS);      // [RES_MAXCOLS][2] sites whose C_o rows scale them
    int *srow = dCol + 2 * RES_MAXCOLS;                    // [RSLOTS]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int4 ct = ctaTab[blockIdx.x];              // tile, first local column, number of columns
    const int4 ti = tileInfo[ct.x];                  // slotBegin, nSlots, first strip-table entry, nStrips
    const int pl0 = ct.y, ncb = ct.z;

    if (tid < RSLOTS) srow[tid] = tid < ti.y ? slotRowG[ti.x + tid] : -1;
    for (int i = tid; i < 3 * RBUF / 2; i += CNT) *reinterpret_cast<double2 *>(Ys + 2 * i) = make_double2(0.0, 0.0);
    // column descriptions: panels of the (<= 2) elements and the C_o rows that scale them
    for (int cb = tid; cb < ncb; cb += CNT) {
        const int p = v.colBegin + pl0 + cb;
        const int eb = v.h1start[p], n = v.h1start[p + 1] - eb;
        dPan[2 * cb] = n > 0 ? (long long)elU[eb] * v.S * v.ldo : -1;
        dCol[2 * cb] = n > 0 ? v.h1col[eb] : 0;
        dPan[2 * cb + 1] = n > 1 ? (long long)elU[eb + 1] * v.S * v.ldo : -1;
        dCol[2 * cb + 1] = n > 1 ? v.h1col[eb + 1] : 0;
    }
    int sw[CMAXS];
#pragma unroll
    for (int s = 0; s < CMAXS; ++s) sw[s] = stripsG[ti.z + warp * CMAXS + s];
    double acc[CMAXS][CMAXN2][2];
#pragma unroll
    for (int s = 0; s < CMAXS; ++s)
#pragma unroll
        for (int j = 0; j < CMAXN2; ++j) acc[s][j][0] = acc[s][j][1] = 0.0;
    __syncthreads();

    // this thread's items: (slot, orbital pair), consecutive threads on consecutive addresses of a panel row
    int goff[NIT], soff[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        const int i = tid + CNT * k, slot = i / QP, q = i % QP;
        const int row = srow[slot];
        goff[k] = row >= 0 ? row * v.ldo + 2 * q : -1;         // S ldo < 2^31 (checked at plan creation)
        soff[k] = slot * RLD + 2 * q;
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k)
        if (goff[k] >= 0) *reinterpret_cast<double2 *>(Ps + soff[k]) = ld2(v.PO + goff[k]);

    const int g4 = lane >> 2, t4 = lane & 3;
    int offA[CMAXS], offB[CMAXS];
#pragma unroll
    for (int s = 0; s < CMAXS; ++s) {
        offA[s] = ((sw[s] & 0xFF) * CB8 + g4) * RLD + t4;
        offB[s] = (((sw[s] >> 8) & 0xFF) * CB8 + g4) * RLD + t4;
    }
    const int nk = (min(LDO, v.no) + 3) / 4;

    // where this warp's strips store: fragment geometry is fixed for the CTA
    int stRs[CMAXS], stRf[CMAXS], stRa[CMAXS];
    long long stQ[CMAXS];
#pragma unroll
    for (int s = 0; s < CMAXS; ++s) {
        stRs[s] = stRf[s] = 0; stRa[s] = -1; stQ[s] = 0;
        if (sw[s] < 0) continue;
        const int b1 = sw[s] & 0xFF;
        const int f = v.rowFrag[srow[b1 * CB8]];                  // the first row of a block always exists
        stRs[s] = v.rowStart[f];
        stRf[s] = v.rowStart[f + 1] - stRs[s];
        stRa[s] = srow[b1 * CB8 + g4];
        stQ[s] = v.qoff[f];
    }
    long long resident = -1;                         // panel offset held in Mr
    long long pA = dPan[0], pB = dPan[1];
    double2 mreg[NIT];                               // streamed panel (second element) of the current column
#pragma unroll
    for (int k = 0; k < NIT; ++k) mreg[k] = (pB >= 0 && goff[k] >= 0) ? ld2(M + pB + goff[k]) : make_double2(0.0, 0.0);
    if (tid < LDO) {
        coS[tid] = pA >= 0 ? v.CoT[(size_t)dCol[0] * v.ldo + tid] : 0.0;
        coS[LDO + tid] = pB >= 0 ? v.CoT[(size_t)dCol[1] * v.ldo + tid] : 0.0;
    }

    for (int cb = 0; cb < ncb; ++cb) {
        const int buf = cb & 1;
        // (A) the first element's panel: reload only when the row site changes
        if (pA >= 0 && pA != resident) {
#pragma unroll
            for (int k = 0; k < NIT; ++k)
                if (goff[k] >= 0) *reinterpret_cast<double2 *>(Mr + soff[k]) = ld2(M + pA + goff[k]);
            resident = pA;
        }
        __syncthreads();                             // coS[buf] visible; every warp is done with the previous Ys
        // (B) Y = M_a * Co[c_a] + M_b * Co[c_b] for this thread's own items (padding slots: 0 = 0 a + 0 b, no branch)
#pragma unroll
        for (int k = 0; k < NIT; ++k) {
            const int q2 = soff[k] % RLD;            // 2 q
            const double2 a = ld2(coS + (2 * buf) * LDO + q2), b = ld2(coS + (2 * buf + 1) * LDO + q2);
            const double2 mr = ld2(Mr + soff[k]);
            double2 y;
            y.x = fma(mreg[k].x, b.x, mr.x * a.x);
            y.y = fma(mreg[k].y, b.y, mr.y * a.y);
            *reinterpret_cast<double2 *>(Ys + soff[k]) = y;
        }
        __syncthreads();                             // Ys complete
        // (C) next column: streamed panel into registers, C_o rows into the other buffer -- in flight during (D)
        long long nA = -1, nB = -1;
        if (cb + 1 < ncb) {
            nA = dPan[2 * cb + 2]; nB = dPan[2 * cb + 3];
#pragma unroll
            for (int k = 0; k < NIT; ++k)
                mreg[k] = (nB >= 0 && goff[k] >= 0) ? ld2(M + nB + goff[k]) : make_double2(0.0, 0.0);
            if (tid < LDO) {
                coS[(2 * (buf ^ 1)) * LDO + tid] = nA >= 0 ? v.CoT[(size_t)dCol[2 * cb + 2] * v.ldo + tid] : 0.0;
                coS[(2 * (buf ^ 1) + 1) * LDO + tid] = nB >= 0 ? v.CoT[(size_t)dCol[2 * cb + 3] * v.ldo + tid] : 0.0;
            }
        }
        // (D) the strips of this warp over the whole occupied range
#pragma unroll
        for (int s = 0; s < CMAXS; ++s) {
            if (sw[s] < 0) continue;                 // warp-uniform
            const double *ya = Ys + offA[s], *pa = Ps + offA[s], *yb = Ys + offB[s], *pb = Ps + offB[s];
            switch (sw[s] >> 16) {                   // warp-uniform
            case 4: strip_steps<4, RLD, true, 8>(acc[s], ya, pa, yb, pb, nk); break;
            case 3: strip_steps<3, RLD, true, 8>(acc[s], ya, pa, yb, pb, nk); break;
            case 2: strip_steps<2, RLD, true, 8>(acc[s], ya, pa, yb, pb, nk); break;
            default: strip_steps<1, RLD, true, 8>(acc[s], ya, pa, yb, pb, nk); break;
            }
        }
        // (E) J[a,b] and J[b,a] from the same register, then restart the sums
        double *Jc = v.J + (size_t)v.Q * (pl0 + cb);
#pragma unroll
        for (int s = 0; s < CMAXS; ++s) {
            if (sw[s] < 0) continue;
            const int b1 = sw[s] & 0xFF, b2 = (sw[s] >> 8) & 0xFF, n2 = sw[s] >> 16;
            const int rs = stRs[s], rf = stRf[s], ra = stRa[s];
            double *Jp = Jc + stQ[s];
#pragma unroll
            for (int j = 0; j < CMAXN2; ++j) {
                if (j < n2) {
                    const bool diag = (b2 + j == b1);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int rb = srow[(b2 + j) * CB8 + 2 * t4 + e];
                        if (ra >= 0 && rb >= 0 && !(diag && 2 * t4 + e < g4)) {
                            const long long al = ra - rs, bl = rb - rs;
                            const double val = 2.0 * acc[s][j][e];
                            QC_DEV_ASSERT(al >= 0 && bl >= 0 && al < rf && bl < rf);
                            Jp[al + rf * bl] = val;
                            Jp[bl + rf * al] = val;
                        }
                    }
                }
                acc[s][j][0] = acc[s][j][1] = 0.0;
            }
        }
        pA = nA; pB = nB;
    }
}

// Warp-specialised variant, OFF by default (QCDMET_B200_CONTRACT_WS=1 selects it).  Measured once at the end
// of round 1: bit-identical to contract_panels_kernel (tests -k experimental) and 2 % faster on the sn2-like
// shape -- the contraction is bound by L1TEX wavefronts, not by exposed latency -- so the
// simpler kernel stays the default.  Same schedule, same arithmetic per accumulator and the same stores; the difference is
// who waits for whom: warps 0-3 only stage (Y, PO chunks into two shared-memory stages, 8 tasks x 3 LDG.128
// in flight per thread), warps 4-11 only run their strips, and the two groups meet on full/empty mbarriers
// per stage instead of CTA-wide barriers, so the L2 latency of the panels overlaps the DMMA work of the
// previous chunk and a slow strip does not hold the other DMMA warps.
template <bool NNZ2>
__global__ void __launch_bounds__(WS_NT, 1)
contract_panels_ws_kernel(const JacView v, const double *__restrict__ M, const int *__restrict__ elU,
                          const int4 *__restrict__ tileInfo, const int *__restrict__ slotRowG,
                          const int *__restrict__ stripsG)
{
    extern __shared__ __align__(16) unsigned char csm[];
    double *Ys = reinterpret_cast<double *>(csm);                    // [stage][slot][o]: Ys, Ps of stage 0, then of stage 1
    long long *elM = reinterpret_cast<long long *>(Ys + 4 * CBUF);   // [CCB][CMAXE]
    long long *elC = elM + CCB * CMAXE;
    int *nnzs = reinterpret_cast<int *>(elC + CCB * CMAXE);
    int *srow = nnzs + CCB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(srow + CSLOTS);    // full[0], full[1], empty[0], empty[1]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int4 ti = tileInfo[blockIdx.x];
    const int pl0 = blockIdx.y * CCB;
    const int ncb = min(CCB, v.colEnd - v.colBegin - pl0);

    if (tid == 0) {
        mbar_init(smem_addr(&bars[0]), WS_STAGE_WARPS);   // full: one arrival per staging warp
        mbar_init(smem_addr(&bars[1]), WS_STAGE_WARPS);
        mbar_init(smem_addr(&bars[2]), CWARPS);           // empty: one arrival per DMMA warp
        mbar_init(smem_addr(&bars[3]), CWARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < CSLOTS) srow[tid] = tid < ti.y ? slotRowG[ti.x + tid] : -1;
    for (int i = tid; i < ncb * CMAXE; i += WS_NT) {
        const int cb = i / CMAXE, e = i % CMAXE;
        const int p = v.colBegin + pl0 + cb;
        const int eb = v.h1start[p], n = v.h1start[p + 1] - eb;
        if (e == 0) nnzs[cb] = n;
        if (e < n) {
            elM[i] = (long long)elU[eb + e] * v.S * v.ldo;
            elC[i] = (long long)v.h1col[eb + e] * v.ldo;
        }
    }
    // padding slots of ragged blocks stay zero in both stages: staging never touches them
    for (int i = tid; i < 2 * CBUF; i += WS_NT) *reinterpret_cast<double2 *>(Ys + 2 * i) = make_double2(0.0, 0.0);
    __syncthreads();

    const int nChunks = (v.no + CTO - 1) / CTO;
    const int nIt = ncb * nChunks;

    if (warp < WS_STAGE_WARPS) {
        // ================= staging warps =================
        // task = (4 consecutive slots, 16-orbital half of the chunk); lane = (slot of the group, orbital pair).
        // Warp w serves half w & 1 of the groups (w >> 1) + 2 i, i < 16.
        const int ls = lane & 3, lo = lane >> 2;
        const int half = warp & 1;
        const int colo = 16 * half + 2 * lo;
        const int dst0 = (4 * (warp >> 1) + ls) * CLD + colo;
        int roff[WS_TASKS];
#pragma unroll
        for (int i = 0; i < WS_TASKS; ++i) {
            const int slot = 4 * ((warp >> 1) + 2 * i) + ls;
            const int row = slot < ti.y ? srow[slot] : -1;
            roff[i] = row >= 0 ? row * v.ldo + colo : -1;
        }
        int cb = 0, ch = 0;
        for (int it = 0; it < nIt; ++it) {
            const int st = it & 1, o0 = ch * CTO;
            double *ys = Ys + 2 * st * CBUF, *ps = ys + CBUF;
            mbar_wait(smem_addr(&bars[2 + st]), ((it >> 1) & 1) ^ 1);      // the DMMA warps are done with this stage
            if (o0 + 16 * half < v.ldo) {
                const int nnz = nnzs[cb];
                const long long *eM = elM + cb * CMAXE, *eC = elC + cb * CMAXE;
                const double *pM0 = M + (nnz > 0 ? eM[0] : 0), *pM1 = M + (nnz > 1 ? eM[1] : 0);
                double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
                if (nnz > 0) c0 = ld2(v.CoT + eC[0] + o0 + colo);
                if (nnz > 1) c1 = ld2(v.CoT + eC[1] + o0 + colo);
#pragma unroll
                for (int i0 = 0; i0 < WS_TASKS; i0 += WS_BATCH) {
                    double2 m0[WS_BATCH], m1[WS_BATCH], po[WS_BATCH];
#pragma unroll
                    for (int k = 0; k < WS_BATCH; ++k) {
                        m0[k] = m1[k] = po[k] = make_double2(0.0, 0.0);
                        if (roff[i0 + k] >= 0) {
                            po[k] = ld2(v.PO + roff[i0 + k] + o0);
                            if (nnz > 0) m0[k] = ld2(pM0 + roff[i0 + k] + o0);
                            if (nnz > 1) m1[k] = ld2(pM1 + roff[i0 + k] + o0);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < WS_BATCH; ++k) {
                        if (roff[i0 + k] < 0) continue;
                        double2 y;
                        y.x = fma(m1[k].x, c1.x, m0[k].x * c0.x);
                        y.y = fma(m1[k].y, c1.y, m0[k].y * c0.y);
                        if constexpr (!NNZ2) {
                            for (int e = 2; e < nnz; ++e) {
                                const double2 m = ld2(M + eM[e] + roff[i0 + k] + o0), c = ld2(v.CoT + eC[e] + o0 + colo);
                                y.x = fma(m.x, c.x, y.x); y.y = fma(m.y, c.y, y.y);
                            }
                        }
                        *reinterpret_cast<double2 *>(ys + dst0 + (i0 + k) * 8 * CLD) = y;
                        *reinterpret_cast<double2 *>(ps + dst0 + (i0 + k) * 8 * CLD) = po[k];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_addr(&bars[st]));
            if (++ch == nChunks) { ch = 0; ++cb; }
        }
        return;
    }

    // ================= DMMA warps =================
    const int dw = warp - WS_STAGE_WARPS;
    const int g4 = lane >> 2, t4 = lane & 3;
    int sw[CMAXS], offA[CMAXS], offB[CMAXS];
#pragma unroll
    for (int s = 0; s < CMAXS; ++s) {
        sw[s] = stripsG[ti.z + dw * CMAXS + s];
        offA[s] = ((sw[s] & 0xFF) * CB8 + g4) * CLD + t4;
        offB[s] = (((sw[s] >> 8) & 0xFF) * CB8 + g4) * CLD + t4;
    }
    double acc[CMAXS][CMAXN2][2];
#pragma unroll
    for (int s = 0; s < CMAXS; ++s)
#pragma unroll
        for (int j = 0; j < CMAXN2; ++j) acc[s][j][0] = acc[s][j][1] = 0.0;
    int cb = 0, ch = 0;
    for (int it = 0; it < nIt; ++it) {
        const int st = it & 1;
        const double *ys = Ys + 2 * st * CBUF, *ps = ys + CBUF;
        mbar_wait(smem_addr(&bars[st]), (it >> 1) & 1);                    // the chunk is staged
        const int nk = (min(CTO, v.no - ch * CTO) + 3) / 4;
#pragma unroll
        for (int s = 0; s < CMAXS; ++s) {
            if (sw[s] < 0) continue;
            const double *ya = ys + offA[s], *pa = ps + offA[s], *yb = ys + offB[s], *pb = ps + offB[s];
            switch (sw[s] >> 16) {
            case 4: strip_steps<4>(acc[s], ya, pa, yb, pb, nk); break;
            case 3: strip_steps<3>(acc[s], ya, pa, yb, pb, nk); break;
            case 2: strip_steps<2>(acc[s], ya, pa, yb, pb, nk); break;
            default: strip_steps<1>(acc[s], ya, pa, yb, pb, nk); break;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_addr(&bars[2 + st]));              // this warp no longer reads the stage
        if (ch == nChunks - 1) {
            double *Jc = v.J + (size_t)v.Q * (pl0 + cb);
#pragma unroll
            for (int s = 0; s < CMAXS; ++s) {
                if (sw[s] < 0) continue;
                const int b1 = sw[s] & 0xFF, b2 = (sw[s] >> 8) & 0xFF, n2 = sw[s] >> 16;
                const int f = v.rowFrag[srow[b1 * CB8]];
                const int rs = v.rowStart[f], rf = v.rowStart[f + 1] - rs;
                const int ra = srow[b1 * CB8 + g4];
                double *Jp = Jc + v.qoff[f];
#pragma unroll
                for (int j = 0; j < CMAXN2; ++j) {
                    if (j < n2) {
                        const bool diag = (b2 + j == b1);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int rb = srow[(b2 + j) * CB8 + 2 * t4 + e];
                            if (ra >= 0 && rb >= 0 && !(diag && 2 * t4 + e < g4)) {
                                const long long al = ra - rs, bl = rb - rs;
                                const double val = 2.0 * acc[s][j][e];
                                QC_DEV_ASSERT(al >= 0 && bl >= 0 && al < rf && bl < rf);
                                Jp[al + rf * bl] = val;
                                Jp[bl + rf * al] = val;
                            }
                        }
                    }
                    acc[s][j][0] = acc[s][j][1] = 0.0;
                }
            }
        }
        if (++ch == nChunks) { ch = 0; ++cb; }
    }
}

// ------------------------------------------------------------------------------------------
// host: plan, launch
// ------------------------------------------------------------------------------------------
int contract_plan_build(ContractPlan **out, int F, const int *rowStart)
{
    *out = nullptr;
    ContractPlan *c = new ContractPlan();
    for (int f = 0; f < F; ++f) c->maxFragBlocks = std::max(c->maxFragBlocks, (rowStart[f + 1] - rowStart[f] + CB8 - 1) / CB8);
    contract_schedule_build(F, rowStart, &c->sched, CMAXB);
    contract_schedule_build(F, rowStart, &c->sched8, RMAXB);
    {
        const ContractSchedule &r = c->sched8;
        const size_t nt8 = std::max<size_t>(r.tileInfo.size(), 4), nsl8 = std::max<size_t>(r.slotRow.size(), 1),
                     ni8 = std::max<size_t>(r.strips.size(), 1);
        if (cudaMalloc((void **)&c->d_tileInfo8, nt8 * sizeof(int)) != cudaSuccess ||
            cudaMalloc((void **)&c->d_slotRow8, nsl8 * sizeof(int)) != cudaSuccess ||
            cudaMalloc((void **)&c->d_strips8, ni8 * sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaMalloc of the contraction schedule failed");
            contract_plan_free(c);
            return QCDMET_E_ALLOC;
        }
        if (!r.tileInfo.empty())
            QC_CUDA(cudaMemcpy(c->d_tileInfo8, r.tileInfo.data(), r.tileInfo.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (!r.slotRow.empty())
            QC_CUDA(cudaMemcpy(c->d_slotRow8, r.slotRow.data(), r.slotRow.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (!r.strips.empty())
            QC_CUDA(cudaMemcpy(c->d_strips8, r.strips.data(), r.strips.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    const ContractSchedule &s = c->sched;
    const size_t nt = std::max<size_t>(s.tileInfo.size(), 4), nsl = std::max<size_t>(s.slotRow.size(), 1),
                 ni = std::max<size_t>(s.strips.size(), 1);
    if (cudaMalloc((void **)&c->d_tileInfo, nt * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&c->d_slotRow, nsl * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&c->d_strips, ni * sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of the contraction schedule failed");
        contract_plan_free(c);
        return QCDMET_E_ALLOC;
    }
    if (!s.tileInfo.empty())
        QC_CUDA(cudaMemcpy(c->d_tileInfo, s.tileInfo.data(), s.tileInfo.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!s.slotRow.empty())
        QC_CUDA(cudaMemcpy(c->d_slotRow, s.slotRow.data(), s.slotRow.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!s.strips.empty())
        QC_CUDA(cudaMemcpy(c->d_strips, s.strips.data(), s.strips.size() * sizeof(int), cudaMemcpyHostToDevice));
    *out = c;
    return QCDMET_OK;
}

void contract_plan_free(ContractPlan *c)
{
    if (!c) return;
    cudaFree(c->d_tileInfo); cudaFree(c->d_slotRow); cudaFree(c->d_strips); cudaFree(c->d_ctaTab);
    cudaFree(c->d_tileInfo8); cudaFree(c->d_slotRow8); cudaFree(c->d_strips8);
    cudaGetLastError();
    delete c;
}

namespace {
// CTAs of the resident variant: every tile gets a number of CTAs proportional to its block-pair count (its DMMA
// work per column), each CTA a run of consecutive columns; about 3 CTAs per SM so that the hardware scheduler evens
// out what is left.  Returns the table (tile, first local column, columns, 0).
std::vector<int4> resident_cta_table(const ContractSchedule &s, int ncols, int numSMs, int ctasPerSM)
{
    // time of a CTA per column: the DMMA phase lasts as long as the busiest warp's strips (block pairs), plus a
    // constant for staging Y, the two CTA-wide barriers and the stores (about 3 block pairs' worth)
    std::vector<double> cost(s.nTiles, 0.0);
    double total = 0.0;
    for (int t = 0; t < s.nTiles; ++t) {
        const int first = s.tileInfo[4 * t + 2];
        int worst = 0;
        for (int w = 0; w < CWARPS; ++w) {
            int load = 0;
            for (int k = 0; k < CMAXS; ++k) {
                const int st = s.strips[first + w * CMAXS + k];
                if (st >= 0) load += st >> 16;
            }
            worst = std::max(worst, load);
        }
        cost[t] = worst + 3.0;
        total += cost[t];
    }
    // CTAs per tile: largest-remainder split of exactly `waves` x numSMs CTAs (no straggler wave)
    const int waves = 3;                          // per resident CTA slot
    long long want = (long long)ctasPerSM * waves * numSMs;
    want = std::max<long long>(want, (long long)s.nTiles * ((ncols + RES_MAXCOLS - 1) / RES_MAXCOLS));
    std::vector<long long> n(s.nTiles, 0);
    std::vector<std::pair<double, int>> rem;
    long long given = 0;
    for (int t = 0; t < s.nTiles; ++t) {
        const double share = want * cost[t] / std::max(total, 1e-30);
        n[t] = std::max<long long>(1, (long long)share);
        given += n[t];
        rem.push_back({share - (double)(long long)share, t});
    }
    std::sort(rem.begin(), rem.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) {
        return a.first > b.first || (a.first == b.first && a.second < b.second);
    });
    for (size_t k = 0; given < want && k < rem.size(); ++k, ++given) ++n[rem[k].second];
    std::vector<int4> tab;
    for (int t = 0; t < s.nTiles; ++t) {
        n[t] = std::min<long long>(n[t], ncols);
        n[t] = std::max<long long>(n[t], (ncols + RES_MAXCOLS - 1) / RES_MAXCOLS);
        for (long long k = 0; k < n[t]; ++k) {
            const int c0 = (int)((long long)ncols * k / n[t]), c1 = (int)((long long)ncols * (k + 1) / n[t]);
            if (c1 > c0) tab.push_back(make_int4(t, c0, c1 - c0, 0));
        }
    }
    return tab;
}

template <int LDO, int RSLOTS>
constexpr int resident_smem()
{
    return 3 * RSLOTS * (LDO + 4) * 8 + 4 * LDO * 8 + RES_MAXCOLS * (2 * 8 + 2 * 4) + RSLOTS * 4;
}
}  // namespace

int launch_contract_panels(ContractPlan *c, const JacView &v, const double *M, const int *elU,
                           long long nnzInRange, bool nnzAtMost2, cudaStream_t st, int *launches, double *flops)
{
    *launches = 0;
    *flops = 0.0;
    const int ncols = v.colEnd - v.colBegin;
    const ContractSchedule &s = c->sched;
    if (ncols <= 0 || s.nTiles == 0) return QCDMET_OK;
    // executed: 2 DMMA.8x8x4 (2 x 512 flop) per block pair and k-step of 4 orbitals (padding of ragged
    // blocks and of no to a multiple of 4 included); Y build 2 flop per element, slot and orbital
    *flops = (double)ncols * (double)s.blockPairs * 1024.0 * ((v.no + 3) / 4) +
             2.0 * (double)nnzInRange * (double)s.validSlots * v.no;
    const char *wsEnv = getenv("QCDMET_B200_CONTRACT_WS");
    const bool ws = wsEnv && wsEnv[0] == '1';
    const char *resEnv = getenv("QCDMET_B200_CONTRACT_RESIDENT");          // "0": keep the chunked kernel (tests, tuning)
    if (nnzAtMost2 && v.ldo <= 64 && !ws && !(resEnv && resEnv[0] == '0')) {
        // ---- resident variant: the whole occupied range of a tile stays in shared memory
        // tiles of 64 row slots with two CTAs per SM when the fragments are small (<= 2 blocks: the strips of a 64-slot
        // tile then still fill the eight warps); tiles of 128 slots, one CTA per SM otherwise
        const bool small = c->maxFragBlocks <= 2;
        const char *slotsEnv = getenv("QCDMET_B200_CONTRACT_SLOTS");        // tuning: 64 | 128
        const bool use64 = slotsEnv ? atoi(slotsEnv) == 64 : small;
        const ContractSchedule &rs = use64 ? c->sched8 : c->sched;
        if (c->resCols != ncols || c->res64 != use64 || !c->d_ctaTab) {
            int dev = 0, sms = 0;
            QC_CUDA(cudaGetDevice(&dev));
            QC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            const std::vector<int4> tab = resident_cta_table(rs, ncols, sms, use64 ? 2 : 1);
            if (tab.size() > c->ctaCapacity) {
                cudaFree(c->d_ctaTab);
                c->d_ctaTab = nullptr;
                c->ctaCapacity = 0;
                if (cudaMalloc((void **)&c->d_ctaTab, tab.size() * sizeof(int4)) != cudaSuccess) {
                    cudaGetLastError();
                    set_error("cudaMalloc of the contraction CTA table failed");
                    return QCDMET_E_ALLOC;
                }
                c->ctaCapacity = tab.size();
            }
            QC_CUDA(cudaMemcpyAsync(c->d_ctaTab, tab.data(), tab.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
            QC_CUDA(cudaStreamSynchronize(st));          // tab is a local: the copy must have read it
            c->nCtas = (int)tab.size();
            c->resCols = ncols;
            c->res64 = use64;
        }
        const int4 *ctaTab = reinterpret_cast<const int4 *>(c->d_ctaTab);
        const int4 *tinfo = reinterpret_cast<const int4 *>(use64 ? c->d_tileInfo8 : c->d_tileInfo);
        const int *slotRow = use64 ? c->d_slotRow8 : c->d_slotRow, *strips = use64 ? c->d_strips8 : c->d_strips;
#define QC_LAUNCH_RESIDENT(LDOV, SLOTSV)                                                                      \
        if (v.ldo == LDOV && (use64 ? 64 : 128) == SLOTSV) {                                                  \
            QC_CUDA(cudaFuncSetAttribute(contract_resident_kernel<LDOV, SLOTSV>,                              \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, resident_smem<LDOV, SLOTSV>())); \
            contract_resident_kernel<LDOV, SLOTSV><<<c->nCtas, CNT, resident_smem<LDOV, SLOTSV>(), st>>>(     \
                v, M, elU, tinfo, slotRow, strips, ctaTab);                                                   \
        }
        QC_LAUNCH_RESIDENT(16, 64) QC_LAUNCH_RESIDENT(32, 64) QC_LAUNCH_RESIDENT(48, 64) QC_LAUNCH_RESIDENT(64, 64)
        QC_LAUNCH_RESIDENT(16, 128) QC_LAUNCH_RESIDENT(32, 128) QC_LAUNCH_RESIDENT(48, 128) QC_LAUNCH_RESIDENT(64, 128)
#undef QC_LAUNCH_RESIDENT
        QC_CUDA(cudaGetLastError());
        *launches += 1;
        return QCDMET_OK;
    }
    // per device (function attributes live in the context of the current device): set on every launch
    QC_CUDA(cudaFuncSetAttribute(contract_panels_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CSMEM));
    QC_CUDA(cudaFuncSetAttribute(contract_panels_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CSMEM));
    // warp-specialised variant (see the kernel's comment): only on request
    if (ws) {
        QC_CUDA(cudaFuncSetAttribute(contract_panels_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
        QC_CUDA(cudaFuncSetAttribute(contract_panels_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
    }
    // gridDim.y is limited to 65535 batches of CCB columns: chunk the columns
    const int chunk = 65535 * CCB;
    for (int c0 = 0; c0 < ncols; c0 += chunk) {
        const int nc = std::min(chunk, ncols - c0);
        JacView w = v;
        w.colBegin = v.colBegin + c0;
        w.colEnd = w.colBegin + nc;
        w.J = v.J + (size_t)v.Q * c0;
        const dim3 grid((unsigned)s.nTiles, (unsigned)((nc + CCB - 1) / CCB));
        if (ws && nnzAtMost2)
            contract_panels_ws_kernel<true><<<grid, WS_NT, WS_SMEM, st>>>(
                w, M, elU, reinterpret_cast<const int4 *>(c->d_tileInfo), c->d_slotRow, c->d_strips);
        else if (ws)
            contract_panels_ws_kernel<false><<<grid, WS_NT, WS_SMEM, st>>>(
                w, M, elU, reinterpret_cast<const int4 *>(c->d_tileInfo), c->d_slotRow, c->d_strips);
        else if (nnzAtMost2)
            contract_panels_kernel<true><<<grid, CNT, CSMEM, st>>>(
                w, M, elU, reinterpret_cast<const int4 *>(c->d_tileInfo), c->d_slotRow, c->d_strips);
        else
            contract_panels_kernel<false><<<grid, CNT, CSMEM, st>>>(
                w, M, elU, reinterpret_cast<const int4 *>(c->d_tileInfo), c->d_slotRow, c->d_strips);
        QC_CUDA(cudaGetLastError());
        *launches += 1;
    }
    return QCDMET_OK;
}

}  // namespace qcdmet
