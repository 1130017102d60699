"""CPU tests of the host-side mirror of qcdmethelper / dmet (no device needed)."""
import numpy as np

from oracle import response_oracle as ro
from qcdmet_b200 import dmet as dmet_mod, qcdmethelper as helper_mod, workloads as wl


class Ints:
    """Minimal duck-type of theLocalIntegrals for this path."""
    TI_OK = True
    ERIinMEM = False

    def __init__(self, fock, nelec):
        self._f = np.asarray(fock)
        self.Norbs = self._f.shape[0]
        self.Nelec = nelec

    def loc_rhf_fock(self):
        return self._f

    def loc_oei(self):
        return self._f


def make_dmet(w, **kw):
    return dmet_mod.dmet(Ints(w.h0, 2 * w.num_pairs), w.imp_clust, w.trans_inv, fitImpBath=w.fit_imp_bath, **kw)


def test_dmet_parameter_lists_match_oracle():
    for w in (wl.sweep(16), wl.make_workload("m", 20, 8, [3, 5, 2]), wl.hydrogen_like(),
              wl.hubbard24(True, l=4)):
        d = make_dmet(w)
        assert (d.helper.H1start == w.h1start).all()
        assert (d.helper.H1row == w.h1row).all() and (d.helper.H1col == w.h1col).all()
        imp_size = ro.make_imp_size(w.imp_clust, w.trans_inv)
        assert (d.imp_size == imp_size).all()
        assert (d.mask == ro.make_mask(w.norb, imp_size)).all()
        x = np.linspace(-1, 1, d.helper.Nterms)
        assert np.array_equal(d.flat2square(x), ro.flat2square(x, w.norb, d.mask, imp_size, w.trans_inv))
        assert np.allclose(d.square2flat(d.flat2square(x)), x)


def test_helper_accepts_dense_and_sparse_lists():
    w = wl.make_workload("m", 12, 5, [3, 4])
    imp_size = ro.make_imp_size(w.imp_clust, False)
    dense = ro.makelist_h1(12, imp_size, False)
    h_dense = helper_mod.qcdmethelper(Ints(w.h0, 10), dense, False, None)
    h_sparse = helper_mod.qcdmethelper(Ints(w.h0, 10), (w.h1start, w.h1row, w.h1col), False, None)
    for a, b in ((h_dense.H1start, h_sparse.H1start), (h_dense.H1row, h_sparse.H1row),
                 (h_dense.H1col, h_sparse.H1col)):
        assert a.dtype == np.int32 and (a == b).all()
    assert h_dense.Nterms == w.nterms and h_dense.numPairs == 5
    assert np.allclose(h_dense.construct1RDM_loc(False, np.zeros((12, 12))),
                       ro.construct1rdm_base(w.h0, 5))


def test_fragment_rows_and_residual():
    w = wl.make_workload("m", 20, 8, [3, 5, 2])
    d = make_dmet(w)
    rs, rows, idx = d._fragment_rows()
    assert rows is None and (rs == w.row_start).all() and (idx == w.row_idx).all()
    rng = np.random.default_rng(0)
    d.imp_1RDM = [rng.standard_normal((2 * m, 2 * m)) for m in (3, 5, 2)]
    x = rng.standard_normal(d.helper.Nterms) * 0.1
    r = d.rdm_differences_host(x)
    r_ref = ro.rdm_differences(w.h0, x, w.num_pairs, w.imp_clust, False, False, d.imp_1RDM)
    assert np.abs(r - r_ref).max() < 1e-13
    wb = wl.make_workload("b", 16, 6, [2, 3], fit_imp_bath=True)
    db = make_dmet(wb)
    db.dmetOrbs = wb.dmet_orbs
    rs, rows, idx = db._fragment_rows()
    assert idx is None and np.array_equal(rows, wb.rows) and (rs == wb.row_start).all()


def test_dmma_schedule_host_logic():
    """Schedule builder of the DMMA kernel (host code inside the shared library; no device needed)."""
    from qcdmet_b200 import _lib

    def info(w):
        return _lib.schedule_info(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.row_idx)

    s = info(wl.sweep(4000))                       # the north-star shape
    assert s["eligible"] and s["symmetric"] and not s["split"]
    assert (s["rows_per_fragment"], s["nj"]) == (4, 6)          # 96-column tiles: no = 2000 -> 2016
    assert s["components"] == 1000 and s["row_tiles"] == 125 and s["column_tiles"] == 21
    assert s["block_pairs"] == 8 * 125 * 126 // 2               # tile(g) >= t: half of 125000 (+ diagonal)
    s = info(wl.sweep(576))
    assert s["nj"] == 6 and s["column_tiles"] == 3              # no = 288 = 3 x 96
    s = info(wl.sweep(160, nterms=333))                         # fragments without perturbations: no mirror mode
    assert s["eligible"] and not s["symmetric"]
    s = info(wl.hubbard24(True))                                # TI: one sub-perturbation per cell
    assert s["eligible"] and s["split"] and s["internal_columns"] == 1440 and s["components"] == 144
    assert s["row_tiles"] == 1 and not s["symmetric"]
    s = info(wl.hubbard24(False, fit_imp_bath=True))            # dense 8-row fragments
    assert s["eligible"] and s["rows_per_fragment"] == 8 and not s["symmetric"] and s["row_tiles"] == 36
    s = info(wl.hydrogen_like())
    assert s["eligible"] and s["split"] and s["rows_per_fragment"] == 4
    s = info(wl.polyyne_like())
    assert not s["eligible"] and "more than" in s["why"]
    s = info(wl.sn2_like())
    assert not s["eligible"]
    mixed = wl.make_workload("mx", 61, 27, [3, 1, 4, 2, 4, 4, 1, 3, 2, 4, 3, 3, 4])
    s = info(mixed)
    assert s["eligible"] and s["symmetric"] and s["components"] == 13 and s["row_tiles"] == 2


def test_constrained_optimisation_variant():
    """use_constrained_opt=True (src/dmet.py:365-379,390-393,441-475): host-only Lagrangian variant.  The
    reference states its gradient -errors "gives correct verify_gradient results": check by finite differences."""
    w = wl.make_workload("alt", 16, 6, [3, 4, 2])
    d = dmet_mod.dmet(Ints(w.h0, 12), w.imp_clust, False, SCmethod='BFGS', fitImpBath=False, use_constrained_opt=True)
    assert d.minFunc == 'FOCK_INIT' and d.helper.altcf
    rng = np.random.default_rng(4)
    d.imp_1RDM = [(lambda a: a + a.T)(rng.standard_normal((2 * m, 2 * m))) for m in (3, 4, 2)]
    x = rng.standard_normal(d.helper.Nterms) * 0.05
    r = d.rdm_differences_bis(x)
    assert r.shape == x.shape
    g = d.alt_costfunction_derivative(x)
    assert np.array_equal(g, -r)
    eps = 1e-6
    for k in (0, 5, len(x) - 1):
        dx = np.zeros_like(x)
        dx[k] = eps
        fd = (d.alt_costfunction(x + dx) - d.alt_costfunction(x - dx)) / (2 * eps)
        # off-diagonal parameters enter umat twice (u_ij and u_ji): the reference's gradient is per parameter
        scale = 1.0 if np.flatnonzero(d.mask.ravel())[k] % 17 == 0 else 2.0
        assert abs(fd - scale * g[k]) < 1e-6 * max(1.0, abs(fd))


def _emulate_contraction(tiles, row_start, y, po):
    """numpy emulation of csrc/kernels_contract.cu for one column, lane by lane: every strip (block b1 against
    blocks b2 .. b2+n2-1) of every warp forms D = Y[b1] PO[b2]^T + PO[b1] Y[b2]^T per 8 x 8 block pair; lane
    (g, t) of the DMMA accumulator holds D[g][2t], D[g][2t+1] and stores 2 D at (a,b) and (b,a)."""
    row_start = np.asarray(row_start)
    qoff = np.concatenate([[0], np.cumsum(np.diff(row_start).astype(np.int64) ** 2)])
    row_frag = np.repeat(np.arange(len(row_start) - 1), np.diff(row_start))
    jcol = np.full(qoff[-1], np.nan)
    writes = np.zeros(qoff[-1], dtype=np.int64)
    for slot_row, per_warp in tiles:
        nslots = len(slot_row)
        assert nslots <= 128 and nslots % 8 == 0 and len(per_warp) == 8
        ys = np.where(slot_row[:, None] >= 0, y[np.maximum(slot_row, 0)], 0.0)     # staged chunk, padding slots zero
        ps = np.where(slot_row[:, None] >= 0, po[np.maximum(slot_row, 0)], 0.0)
        for strips in per_warp:
            assert len(strips) <= 3
            for b1, b2, n2 in strips:
                assert 1 <= n2 <= 4 and b1 <= b2 and 8 * (b2 + n2) <= nslots
                f = row_frag[slot_row[8 * b1]]
                rs, rf = row_start[f], row_start[f + 1] - row_start[f]
                for j in range(n2):
                    bb = b2 + j
                    assert slot_row[8 * bb] >= 0 and row_frag[slot_row[8 * bb]] == f
                    d = ys[8 * b1:8 * b1 + 8] @ ps[8 * bb:8 * bb + 8].T + ps[8 * b1:8 * b1 + 8] @ ys[8 * bb:8 * bb + 8].T
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        ra = slot_row[8 * b1 + g]
                        for e in range(2):
                            rb = slot_row[8 * bb + 2 * t + e]
                            if ra < 0 or rb < 0 or (bb == b1 and 2 * t + e < g):
                                continue
                            al, bl = ra - rs, rb - rs
                            for q in {qoff[f] + al + rf * bl, qoff[f] + bl + rf * al}:
                                jcol[q] = 2.0 * d[g, 2 * t + e]
                                writes[q] += 1
    return jcol, writes


def test_contraction_schedule_covers_every_block_element_once():
    """The host-built tile schedule of the factorised path (qcdmet_contract_schedule): every element of every
    fragment block has exactly one owner, tiles respect the kernel's limits, and the emulated contraction
    equals 2 (G + G^T) block by block -- small packed fragments, ragged blocks, fragments sliced into
    8-block pairs (> 96 rows), empty fragments."""
    from qcdmet_b200 import _lib
    rng = np.random.default_rng(3)
    cases = [[14] * 12, [52] + [24] * 10 + [29], [1, 2, 3, 4, 5, 6, 7, 8, 9], [0, 5, 0, 0, 13, 0], [96], [97], [88, 90, 104],
             [200, 3, 131], [576], [4] * 70, [9] * 40, [], [0, 0]]
    for sizes in cases:
        row_start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        tiles = _lib.contract_schedule(row_start)
        s, no = int(row_start[-1]), 5
        if s == 0:
            assert tiles == []
            continue
        y, po = rng.standard_normal((s, no)), rng.standard_normal((s, no))
        jcol, writes = _emulate_contraction(tiles, row_start, y, po)
        assert writes.min() == 1 and writes.max() == 1
        ref = []
        for a, b in zip(row_start[:-1], row_start[1:]):
            g = y[a:b] @ po[a:b].T
            ref.append((2.0 * (g + g.T)).reshape(-1, order="F"))
        assert np.allclose(jcol, np.concatenate(ref), rtol=0, atol=1e-12)
    # packing: the 12 x 14-row polyyne fragments (2 blocks, 2 strips each) need 2 tiles, the sn2 table 3
    assert len(_lib.contract_schedule(np.arange(0, 14 * 13, 14))) == 2
    assert len(_lib.contract_schedule(np.concatenate([[0], np.cumsum([52] + [24] * 10 + [29])]))) == 3
    # a whole fragment stays in one tile up to 12 blocks (96 rows); beyond that it is cut into slice pairs
    assert len(_lib.contract_schedule(np.array([0, 96]))) == 1
    assert len(_lib.contract_schedule(np.array([0, 97]))) == 3


def test_auto_kernel_selection_rules():
    """Host-only mirror of the selection made at plan creation (qcdmet_auto_kernel): fused DMMA kernel for
    impurities of <= 4 orbitals (<= 8 rows with bath), factorised path for larger fragments with <= 64 elements
    per perturbation, general kernel otherwise."""
    from qcdmet_b200 import _lib
    dmma, fact, general = _lib.KERNEL_DMMA, _lib.KERNEL_FACTORISED, _lib.KERNEL_GENERAL
    cases = [
        (wl.sweep(64), dmma), (wl.hubbard24(False, l=8), dmma), (wl.hubbard24(True, l=8), dmma),
        (wl.hubbard24(True, fit_imp_bath=True, l=8), dmma), (wl.hydrogen_like(), dmma),
        (wl.make_workload("b4", 44, 20, [4, 4], fit_imp_bath=True), dmma),
        (wl.polyyne_like(), fact), (wl.sn2_like(), fact), (wl.make_workload("f5", 40, 17, [5, 4, 3]), fact),
        (wl.make_workload("b5", 44, 20, [5, 3], fit_imp_bath=True), fact),
        (wl.make_workload("ti5", 40, 18, [5] * 8, trans_inv=True), fact),          # 8 cells: 8 / 16 elements
        (wl.make_workload("ti5g", 200, 90, [5] * 40, trans_inv=True), general),    # 40 cells: 80 elements
    ]
    for w, expect in cases:
        kernel, nbytes = _lib.auto_kernel(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.row_idx)
        assert kernel == expect, w.name
        assert (nbytes > 0) == (expect == fact)
    # site panels of the sn2-like shape: 324 site slots x 321 rows x 64 occupied columns x 8 bytes + tiled operands
    _, nbytes = _lib.auto_kernel(321, 58, *(lambda w: (w.h1start, w.h1row, w.h1col, w.row_start, w.row_idx))(wl.sn2_like()))
    assert 324 * 321 * 64 * 8 <= nbytes < 1.1 * 324 * 321 * 64 * 8
    # the compat full response (one fragment holding every site) goes through the factorised path
    w = wl.sweep(100, nterms=23)
    kernel, _ = _lib.auto_kernel(100, 50, w.h1start, w.h1row, w.h1col, np.array([0, 100]), np.arange(100))
    assert kernel == fact
