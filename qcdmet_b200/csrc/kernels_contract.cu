// kernels_contract.cu -- per-perturbation contraction of the site panels M_r (factorised path for
// fragments of more than 8 rows: polyyne- / sn2-like fits and the compat full response).
//
// kernels_dmma.cu (panel mode) leaves  M_r[a,o] = sum_v PV[a,v] Cv[r,v] W[v,o]  in HBM for every site r
// that occurs as an element row.  Per column p (lib/rhf_response.cc:82-108 restricted to fragment blocks):
//   Y_p[a,o]      = sum_{(r,c) in H1_p} M_r[a,o] Co[c,o]
//   J[(f;a,b),p]  = 2 sum_o ( Y_p[a,o] PO[b,o] + PO[a,o] Y_p[b,o] )        a, b rows of fragment f
// FP64 FMA pipe (same issue rate as DMMA on B200).  Work decomposition, built once per plan on the host:
//   unit  = 4 consecutive rows of one fragment (last unit of a fragment may be ragged)
//   item  = unordered pair of units (u1 <= u2) of the same fragment: a 4 x 4 register tile, one thread
//   tile  = <= 32 units (128 row slots) and <= 256 items: whole small fragments packed together, or
//           a pair of 16-unit slices of a large fragment.  One CTA per (tile, column).
// Every (a,b) of every block has exactly one owner and J[a,b], J[b,a] are stored from the same register:
// no atomics, blocks exactly symmetric, run-to-run bit-reproducible.
#include "common.cuh"
#include <algorithm>

namespace qcdmet {

namespace {
constexpr int CU = 4;                  // rows per unit
constexpr int CMAXU = 32;              // units per tile
constexpr int CSLICE = 16;             // units per slice of a large fragment
constexpr int CSLOTS = CU * CMAXU;     // row slots per tile
constexpr int CNT = 256;               // threads per CTA = max items per tile
constexpr int CTO = 16;                // occupied orbitals per staged chunk
constexpr int CCB = 8;                 // columns per CTA (consecutive, pipelined through the same two buffers)
constexpr int CTASK = CMAXU / (CNT / 32);   // staging tasks (units) per warp and chunk: 4
constexpr int CMAXE = QCDMET_CONTRACT_MAX_NNZ;
// Shared-memory layout of a staged chunk: Ys[o][col(u) + i], col(u) = 4 u + 2 (u / 4).  The 16 B pad after
// every 4 units makes the LDS.128 of 8 consecutive units hit 8 distinct 16 B bank groups, and the row
// stride CLD = 2 (mod 16) doubles makes the staging stores (lanes = 8 orbital pairs x 4 rows) conflict-free.
constexpr int CLD = CSLOTS + 2 * (CMAXU / 4) + 2;      // 146
constexpr int CBUF = CTO * CLD;                        // doubles per staged array
constexpr int CSMEM = 4 * CBUF * 8 + CCB * CMAXE * 16 + CCB * 4 + CSLOTS * 4;
__device__ __forceinline__ int unit_col(int u) { return CU * u + 2 * (u >> 2); }
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
}  // namespace

// ------------------------------------------------------------------------------------------
// host: schedule
// ------------------------------------------------------------------------------------------
void contract_schedule_build(int F, const int *rowStart, ContractSchedule *out)
{
    ContractSchedule &s = *out;
    s.tileInfo.clear(); s.slotRow.clear(); s.items.clear();
    s.nTiles = 0; s.validSlots = 0;
    int units = 0, nitems = 0, slotBegin = 0, itemBegin = 0;
    auto flush = [&]() {
        if (units > 0) {
            s.tileInfo.push_back(slotBegin); s.tileInfo.push_back(units * CU);
            s.tileInfo.push_back(itemBegin); s.tileInfo.push_back(nitems);
            ++s.nTiles;
        }
        units = 0; nitems = 0;
        slotBegin = (int)s.slotRow.size(); itemBegin = (int)s.items.size();
    };
    auto push_units = [&](int rs, int rf, int u0, int nu) {
        for (int u = u0; u < u0 + nu; ++u)
            for (int i = 0; i < CU; ++i) {
                const int al = u * CU + i;
                s.slotRow.push_back(al < rf ? rs + al : -1);
                if (al < rf) ++s.validSlots;
            }
    };
    for (int f = 0; f < F; ++f) {
        const int rs = rowStart[f], rf = rowStart[f + 1] - rs;
        if (rf <= 0) continue;
        const int nu = (rf + CU - 1) / CU;
        const int pairs = nu * (nu + 1) / 2;
        if (nu <= CMAXU && pairs <= CNT) {
            // whole fragment in one tile, packed with its neighbours
            if (units + nu > CMAXU || nitems + pairs > CNT) flush();
            const int base = units;
            push_units(rs, rf, 0, nu);
            for (int u1 = 0; u1 < nu; ++u1)
                for (int u2 = u1; u2 < nu; ++u2) s.items.push_back((unsigned short)((base + u1) | ((base + u2) << 8)));
            units += nu; nitems += pairs;
        } else {
            // large fragment: one tile per pair of 16-unit slices (si <= sj)
            flush();
            const int ns = (nu + CSLICE - 1) / CSLICE;
            for (int si = 0; si < ns; ++si)
                for (int sj = si; sj < ns; ++sj) {
                    const int ui0 = si * CSLICE, uin = std::min(CSLICE, nu - ui0);
                    const int uj0 = sj * CSLICE, ujn = std::min(CSLICE, nu - uj0);
                    push_units(rs, rf, ui0, uin);
                    if (sj == si) {
                        for (int u1 = 0; u1 < uin; ++u1)
                            for (int u2 = u1; u2 < uin; ++u2) s.items.push_back((unsigned short)(u1 | (u2 << 8)));
                        units = uin; nitems = uin * (uin + 1) / 2;
                    } else {
                        push_units(rs, rf, uj0, ujn);
                        for (int u1 = 0; u1 < uin; ++u1)
                            for (int u2 = 0; u2 < ujn; ++u2) s.items.push_back((unsigned short)(u1 | ((uin + u2) << 8)));
                        units = uin + ujn; nitems = uin * ujn;
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
// One CTA = (tile, batch of CCB consecutive columns).  The (column, 16-orbital chunk) pairs of the batch run
// through two shared-memory buffers: the global loads of pair it+1 are issued into registers before the
// FMA loop of pair it and committed to the other buffer after it (one barrier per pair), so the L2 latency
// of the panels hides behind the arithmetic and only the first chunk of a CTA is exposed.
__global__ void __launch_bounds__(CNT, 2)
contract_panels_kernel(const JacView v, const double *__restrict__ M, const int *__restrict__ elU,
                       const int4 *__restrict__ tileInfo, const int *__restrict__ slotRowG,
                       const unsigned short *__restrict__ itemsG)
{
    extern __shared__ __align__(16) unsigned char csm[];
    double *Ys = reinterpret_cast<double *>(csm);                    // Ys[buf][o][col]  Y_p of the tile's rows
    double *Ps = Ys + 2 * CBUF;                                      // Ps[buf][o][col]  PO of the same rows
    long long *elM = reinterpret_cast<long long *>(Ps + 2 * CBUF);   // [CCB][CMAXE] offset of the panel M_r of element e
    long long *elC = elM + CCB * CMAXE;                              // [CCB][CMAXE] offset of the row Co[c,:] of element e
    int *nnzs = reinterpret_cast<int *>(elC + CCB * CMAXE);          // [CCB] elements of the column
    int *srow = nnzs + CCB;                                          // [CSLOTS] fragment row of the slot or -1

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int4 ti = tileInfo[blockIdx.x];            // slotBegin, nSlots, itemBegin, nItems
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
    const bool active = tid < ti.w;
    int u1 = 0, u2 = 0;
    if (active) {
        const unsigned it = itemsG[ti.z + tid];
        u1 = it & 0xFF; u2 = it >> 8;
    }
    double acc[CU][CU];
#pragma unroll
    for (int i = 0; i < CU; ++i)
#pragma unroll
        for (int j = 0; j < CU; ++j) acc[i][j] = 0.0;
    __syncthreads();

    // staging geometry of this thread (the same for every chunk): task k = unit warp + 8 k; lane = (row of the
    // unit, orbital pair): a warp load is 4 rows x 128 contiguous bytes, a warp store is conflict-free
    const int ls = lane & 3, lo = lane >> 2;
    const int nUnits = ti.y / CU;
    int src[CTASK], dst[CTASK];          // offset of (row, orbital pair) in a panel, < 0: padding slot; smem offset, < 0: no unit
#pragma unroll
    for (int k = 0; k < CTASK; ++k) {
        const int u = warp + (CNT / 32) * k;
        const int row = u < nUnits ? srow[u * CU + ls] : -1;
        src[k] = row >= 0 ? row * v.ldo + 2 * lo : -1;           // S ldo < 2^31 (checked at plan creation)
        dst[k] = u < nUnits ? 2 * lo * CLD + unit_col(u) + ls : -1;
    }
    const int c1 = unit_col(u1), c2 = unit_col(u2);
    // the first row of a unit always exists; both units of an item belong to the same fragment
    int f = 0, rs = 0, rf = 0;
    if (active) {
        f = v.rowFrag[srow[u1 * CU]];
        rs = v.rowStart[f];
        rf = v.rowStart[f + 1] - rs;
    }

    const int nChunks = (v.no + CTO - 1) / CTO;      // every chunk lies inside the zero-padded ldo (multiple of 16)
    const int nIt = ncb * nChunks;
    double2 m0[CTASK], m1[CTASK], po[CTASK];

    auto issue = [&](int cb, int o0) {               // global loads of one (column, chunk) pair -> registers
        const int nnz = nnzs[cb];
        const long long e0 = nnz > 0 ? elM[cb * CMAXE] : 0, e1 = nnz > 1 ? elM[cb * CMAXE + 1] : 0;
#pragma unroll
        for (int k = 0; k < CTASK; ++k) {
            m0[k] = m1[k] = po[k] = make_double2(0.0, 0.0);
            if (src[k] >= 0) {
                po[k] = ld2(v.PO + src[k] + o0);
                if (nnz > 0) m0[k] = ld2(M + e0 + src[k] + o0);
                if (nnz > 1) m1[k] = ld2(M + e1 + src[k] + o0);
            }
        }
    };
    auto commit = [&](int cb, int o0, int buf) {     // registers -> Y, PO chunk in shared memory
        const int nnz = nnzs[cb];
        const long long *eM = elM + cb * CMAXE, *eC = elC + cb * CMAXE;
        double *ys = Ys + buf * CBUF, *ps = Ps + buf * CBUF;
        const int o = o0 + 2 * lo;
#pragma unroll
        for (int k = 0; k < CTASK; ++k) {
            if (dst[k] < 0) continue;
            double2 y = make_double2(0.0, 0.0);
            if (src[k] >= 0) {
                if (nnz > 0) { const double2 c = ld2(v.CoT + eC[0] + o); y.x = m0[k].x * c.x; y.y = m0[k].y * c.y; }
                if (nnz > 1) { const double2 c = ld2(v.CoT + eC[1] + o); y.x = fma(m1[k].x, c.x, y.x); y.y = fma(m1[k].y, c.y, y.y); }
                for (int e = 2; e < nnz; ++e) {      // longer lists (translation-invariant, arbitrary CSR): not prefetched
                    const double2 m = ld2(M + eM[e] + src[k] + o0), c = ld2(v.CoT + eC[e] + o);
                    y.x = fma(m.x, c.x, y.x); y.y = fma(m.y, c.y, y.y);
                }
            }
            ys[dst[k]] = y.x; ys[dst[k] + CLD] = y.y;
            ps[dst[k]] = po[k].x; ps[dst[k] + CLD] = po[k].y;
        }
    };

    issue(0, 0);
    commit(0, 0, 0);
    __syncthreads();
    int cb = 0, ch = 0;                              // the pair being computed
    for (int it = 0; it < nIt; ++it) {
        const int buf = it & 1;
        int cbn = cb, chn = ch + 1;                  // the pair being fetched
        if (chn == nChunks) { chn = 0; ++cbn; }
        if (it + 1 < nIt) issue(cbn, chn * CTO);
        if (active) {
            const int ow = min(CTO, v.no - ch * CTO);
            const double *y1 = Ys + buf * CBUF + c1, *y2 = Ys + buf * CBUF + c2;
            const double *q1 = Ps + buf * CBUF + c1, *q2 = Ps + buf * CBUF + c2;
#pragma unroll 4
            for (int oo = 0; oo < ow; ++oo) {
                const int off = oo * CLD;
                const double2 ya = ld2(y1 + off), yb = ld2(y1 + off + 2), pa = ld2(q1 + off), pb = ld2(q1 + off + 2);
                const double2 yc = ld2(y2 + off), yd = ld2(y2 + off + 2), pc = ld2(q2 + off), pd = ld2(q2 + off + 2);
                const double Y1[CU] = {ya.x, ya.y, yb.x, yb.y}, P1[CU] = {pa.x, pa.y, pb.x, pb.y};
                const double Y2[CU] = {yc.x, yc.y, yd.x, yd.y}, P2[CU] = {pc.x, pc.y, pd.x, pd.y};
#pragma unroll
                for (int i = 0; i < CU; ++i)
#pragma unroll
                    for (int j = 0; j < CU; ++j) acc[i][j] = fma(Y1[i], P2[j], fma(P1[i], Y2[j], acc[i][j]));
            }
            if (ch == nChunks - 1) {
                // column done: J[a,b] and J[b,a] from the same register, then restart the sums
                double *Jp = v.J + (size_t)v.Q * (pl0 + cb) + v.qoff[f];
#pragma unroll
                for (int i = 0; i < CU; ++i) {
                    const int ra = srow[u1 * CU + i];
#pragma unroll
                    for (int j = 0; j < CU; ++j) {
                        const int rb = srow[u2 * CU + j];
                        if (ra >= 0 && rb >= 0 && !(u1 == u2 && j < i)) {
                            const long long al = ra - rs, bl = rb - rs;
                            const double val = 2.0 * acc[i][j];
                            Jp[al + rf * bl] = val;
                            Jp[bl + rf * al] = val;
                        }
                        acc[i][j] = 0.0;
                    }
                }
            }
        }
        if (it + 1 < nIt) commit(cbn, chn * CTO, buf ^ 1);
        __syncthreads();
        cb = cbn; ch = chn;
    }
}

// ------------------------------------------------------------------------------------------
// host: plan, launch
// ------------------------------------------------------------------------------------------
int contract_plan_build(ContractPlan **out, int F, const int *rowStart)
{
    *out = nullptr;
    ContractPlan *c = new ContractPlan();
    contract_schedule_build(F, rowStart, &c->sched);
    const ContractSchedule &s = c->sched;
    const size_t nt = std::max<size_t>(s.tileInfo.size(), 4), nsl = std::max<size_t>(s.slotRow.size(), 1),
                 ni = std::max<size_t>(s.items.size(), 1);
    if (cudaMalloc((void **)&c->d_tileInfo, nt * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&c->d_slotRow, nsl * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&c->d_items, ni * sizeof(unsigned short)) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of the contraction schedule failed");
        contract_plan_free(c);
        return QCDMET_E_ALLOC;
    }
    if (!s.tileInfo.empty())
        QC_CUDA(cudaMemcpy(c->d_tileInfo, s.tileInfo.data(), s.tileInfo.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!s.slotRow.empty())
        QC_CUDA(cudaMemcpy(c->d_slotRow, s.slotRow.data(), s.slotRow.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!s.items.empty())
        QC_CUDA(cudaMemcpy(c->d_items, s.items.data(), s.items.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    *out = c;
    return QCDMET_OK;
}

void contract_plan_free(ContractPlan *c)
{
    if (!c) return;
    cudaFree(c->d_tileInfo); cudaFree(c->d_slotRow); cudaFree(c->d_items);
    cudaGetLastError();
    delete c;
}

int launch_contract_panels(const ContractPlan *c, const JacView &v, const double *M, const int *elU,
                           long long nnzInRange, cudaStream_t st, int *launches, double *flops)
{
    *launches = 0;
    *flops = 0.0;
    const int ncols = v.colEnd - v.colBegin;
    const ContractSchedule &s = c->sched;
    if (ncols <= 0 || s.nTiles == 0) return QCDMET_OK;
    // per device (function attributes live in the context of the current device): set on every launch
    QC_CUDA(cudaFuncSetAttribute(contract_panels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CSMEM));
    // gridDim.y is limited to 65535 batches of CCB columns: chunk the columns
    const int chunk = 65535 * CCB;
    for (int c0 = 0; c0 < ncols; c0 += chunk) {
        const int nc = std::min(chunk, ncols - c0);
        JacView w = v;
        w.colBegin = v.colBegin + c0;
        w.colEnd = w.colBegin + nc;
        w.J = v.J + (size_t)v.Q * c0;
        contract_panels_kernel<<<dim3((unsigned)s.nTiles, (unsigned)((nc + CCB - 1) / CCB)), CNT, CSMEM, st>>>(
            w, M, elU, reinterpret_cast<const int4 *>(c->d_tileInfo), c->d_slotRow, c->d_items);
        QC_CUDA(cudaGetLastError());
        *launches += 1;
    }
    // executed: 4 FMA-pairs per accumulator element and orbital; Y build 2 flop per element, slot and orbital
    *flops = (double)ncols * (double)s.items.size() * (CU * CU * 4.0) * v.no +
             2.0 * (double)nnzInRange * (double)s.validSlots * v.no;
    return QCDMET_OK;
}

}  // namespace qcdmet
