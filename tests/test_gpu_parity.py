"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the
oracle / the compiled reference's golden outputs.  Tolerance: 1e-10 absolute on response and
Jacobian elements (BASELINE.json north_star)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import response_oracle as ro
from qcdmet_b200 import _lib, dmet as dmet_mod, workloads as wl
from util import GOLDEN, golden_ids, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-10
KERNELS = {"general": _lib.KERNEL_GENERAL, "auto": _lib.KERNEL_AUTO, "factorised": _lib.KERNEL_FACTORISED}
GENERAL, DMMA, FACT = _lib.KERNEL_GENERAL, _lib.KERNEL_DMMA, _lib.KERNEL_FACTORISED


def fact_applies(h1start, nrows):
    """The factorised path stages <= 64 elements per perturbation and needs a non-empty problem."""
    return len(h1start) > 1 and nrows > 0 and int(np.diff(h1start).max()) <= 64


def expect_fact_refused(*plan_args):
    plan = _lib.Plan(*plan_args)
    with pytest.raises(_lib.QcdmetError) as ei:
        plan.set_kernel(FACT)
    assert ei.value.code == -1
    plan.close()


def run_plan(w, kernel, from_orbitals=None, cols=None):
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_kernel(kernel)
    if cols is not None:
        plan.set_columns(*cols)
    if from_orbitals is not None:
        plan.set_orbitals(*from_orbitals)
    else:
        plan.set_h0(w.h0)
    plan.compute()
    jac, c, e, d = plan.jacobian(), plan.orbitals(), plan.eigenvalues(), plan.dblocks()
    used = plan.kernel_used
    run_plan.last_symmetric = plan.symmetric_used
    plan.close()
    return jac, c, e, d, used


def oracle_jac(w, c, e):
    return ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start,
                                     w.rows, w.row_idx)


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_jacobian_matches_golden(path, kernel):
    """End to end from H0 (cuSOLVER syevd on the device) against the compiled reference's output."""
    g = load_golden(path)
    args = (g["norb"], g["num_pairs"], g["h1start"], g["h1row"], g["h1col"], g["row_start"], g["rows"], g["row_idx"])
    if kernel == "factorised" and not fact_applies(g["h1start"], int(g["row_start"][-1])):
        return expect_fact_refused(*args)
    plan = _lib.Plan(*args)
    plan.set_kernel(KERNELS[kernel])
    plan.set_h0(g["h0"])
    plan.compute()
    jac = plan.jacobian()
    assert jac.shape == g["jac"].shape and jac.flags.f_contiguous
    assert np.abs(jac - g["jac"]).max() <= TOL
    plan.close()


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_compat_rhf_response_matches_golden(path):
    """The compat symbol with the reference's own ctypes call shape (src/qcdmethelper.py:102-109)."""
    g = load_golden(path)
    lib = _lib.load()
    n, p = g["norb"], len(g["h1start"]) - 1
    h0 = np.array(g["h0"].reshape(n * n), dtype=ctypes.c_double)
    out = np.ones(n * n * p, dtype=ctypes.c_double)
    rc = lib.rhf_response(ctypes.c_int(n), ctypes.c_int(p), ctypes.c_int(g["num_pairs"]),
                          g["h1start"].ctypes.data_as(ctypes.c_void_p), g["h1row"].ctypes.data_as(ctypes.c_void_p),
                          g["h1col"].ctypes.data_as(ctypes.c_void_p), h0.ctypes.data_as(ctypes.c_void_p),
                          out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, lib.qcdmet_last_error()
    resp = out.reshape((p, n, n), order="C")
    assert np.abs(resp - g["resp"]).max() <= TOL
    assert np.abs(h0.reshape(n, n) - g["one_rdm"]).max() <= TOL      # H0 overwritten by the 1-RDM
    assert np.abs(resp - resp.transpose(0, 2, 1)).max() == 0.0        # exactly symmetric, like the reference


SHAPES = {
    "sweep200": lambda: wl.sweep(200),
    "sweep96_ragged_cols": lambda: wl.sweep(96, nterms=101),
    "hubbard8x8": lambda: wl.hubbard24(False, l=8),
    "hubbard12x12_TI": lambda: wl.hubbard24(True, l=12),
    "hubbard8x8_TI_bath": lambda: wl.hubbard24(True, fit_imp_bath=True, l=8),
    "polyyne_like": wl.polyyne_like,
    "hydrogen_like": wl.hydrogen_like,
    "odd_sizes": lambda: wl.make_workload("odd", 67, 23, [6, 1, 5, 2, 9, 3]),
    "bath_mixed": lambda: wl.make_workload("bath", 50, 21, [3, 4, 2, 5], fit_imp_bath=True),
    "one_occupied": lambda: wl.make_workload("one_occ", 12, 1, [4, 4, 4]),
    "one_virtual": lambda: wl.make_workload("one_virt", 12, 11, [4, 4, 4]),
}


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("shape", list(SHAPES))
def test_jacobian_matches_oracle_on_same_orbitals(shape, kernel):
    w = SHAPES[shape]()
    if kernel == "factorised" and not fact_applies(w.h1start, w.nrows):
        return expect_fact_refused(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    jac, c, e, d, used = run_plan(w, KERNELS[kernel])
    assert kernel == "auto" or used == KERNELS[kernel]
    ref = oracle_jac(w, c, e)
    assert jac.shape == (w.nresid, w.nterms)
    assert np.abs(jac - ref).max() <= TOL
    # eigenvalues / 1-RDM blocks from the device eigensolver agree with LAPACK
    e_ref, c_ref = ro.eigh_sorted(w.h0)
    assert np.abs(e - e_ref).max() < 1e-11
    co = c_ref[:, :w.num_pairs]
    dfull = 2.0 * co @ co.T
    if w.rows is not None:
        blocks = [w.rows[a:b] @ dfull @ w.rows[a:b].T for a, b in zip(w.row_start[:-1], w.row_start[1:])]
    else:
        blocks = [dfull[np.ix_(w.row_idx[a:b], w.row_idx[a:b])] for a, b in zip(w.row_start[:-1], w.row_start[1:])]
    dref = np.concatenate([b.reshape(-1, order="F") for b in blocks])
    assert np.abs(d - dref).max() < 1e-11


DMMA_SHAPES = {
    # name: (workload factory, kernel AUTO must pick, symmetric (mirrored-patch) mode expected)
    "sweep64": (lambda: wl.sweep(64), DMMA, True),
    "sweep264_no_tile_edge": (lambda: wl.sweep(264), DMMA, True),            # no = 132: ragged 128-column tile
    "sweep400": (lambda: wl.sweep(400), DMMA, True),                         # 13 row tiles: many off-diagonal pairs
    "mixed_small_frags": (lambda: wl.make_workload("mx", 61, 27, [3, 1, 4, 2, 4, 4, 1, 3, 2, 4, 3, 3, 4]), DMMA, True),
    "partial_cover": (lambda: wl.make_workload("pc", 90, 40, [4, 4, 2, 3]), DMMA, True),
    "many_tiles": (lambda: wl.make_workload("mt", 150, 70, [3] * 50), DMMA, True),   # 10 + 2/3 fragments per tile
    "tiny_virtual_space": (lambda: wl.make_workload("tv", 40, 37, [4] * 10), DMMA, True),
    "tiny_occupied_space": (lambda: wl.make_workload("to", 40, 2, [4] * 10), DMMA, True),
    "truncated_list": (lambda: wl.sweep(160, nterms=333), DMMA, False),      # fragments without perturbations
    "bath_rows_m2": (lambda: wl.make_workload("b2", 44, 20, [2, 2, 1, 2], fit_imp_bath=True), DMMA, False),
    "frag5_factorised": (lambda: wl.make_workload("f5", 40, 17, [5, 4, 3]), FACT, False),
    "ti_split": (lambda: wl.hubbard24(True, l=8), DMMA, False),              # TI list split per cell, columns summed
    "ti_split_bath": (lambda: wl.hubbard24(True, fit_imp_bath=True, l=8), DMMA, False),
    "ti_m2_bath": (lambda: wl.make_workload("ti2", 24, 11, [2] * 12, trans_inv=True, fit_imp_bath=True), DMMA, False),
    "ti_m3": (lambda: wl.make_workload("ti3", 36, 16, [3] * 12, trans_inv=True), DMMA, False),
    "ti_m5_factorised": (lambda: wl.make_workload("ti5", 40, 18, [5] * 8, trans_inv=True), FACT, False),
    "ti_m5_many_elements_general": (lambda: wl.make_workload("ti5g", 200, 90, [5] * 40, trans_inv=True), GENERAL, False),
    "bath_rows_m4": (lambda: wl.make_workload("b4", 44, 20, [4, 4], fit_imp_bath=True), DMMA, False),   # 8-row fragments
    "bath_rows_mixed": (lambda: wl.make_workload("bm", 70, 31, [4, 3, 1, 4, 2, 4, 4, 3], fit_imp_bath=True), DMMA, False),
    "hubbard8x8_bath": (lambda: wl.hubbard24(False, fit_imp_bath=True, l=8), DMMA, False),
    "sweep192_tile96": (lambda: wl.sweep(192), DMMA, True),                  # no = 96: 96-column tiles
    "frag9_factorised": (lambda: wl.make_workload("f9", 48, 21, [9, 4]), FACT, False),
    "bath_rows_m5_factorised": (lambda: wl.make_workload("b5", 44, 20, [5, 3], fit_imp_bath=True), FACT, False),
}


@pytest.mark.parametrize("shape", list(DMMA_SHAPES))
def test_kernel_selection_and_dmma_parity(shape):
    """AUTO picks the fused DMMA kernel exactly where its schedule applies (<= 4 coupled sites, <= 8 rows
    per fragment), else the factorised path (site panels + contraction; <= 64 elements per perturbation),
    else the general kernel; all match the oracle and each other."""
    factory, expect, expect_sym = DMMA_SHAPES[shape]
    w = factory()
    jac, c, e, d, used = run_plan(w, _lib.KERNEL_AUTO)
    assert used == expect
    assert run_plan.last_symmetric == expect_sym
    assert np.abs(jac - oracle_jac(w, c, e)).max() <= TOL
    gen, *_ = run_plan(w, _lib.KERNEL_GENERAL, from_orbitals=(c, e))
    assert np.abs(jac - gen).max() <= TOL
    if expect != DMMA:
        plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
        with pytest.raises(_lib.QcdmetError):
            plan.set_kernel(_lib.KERNEL_DMMA)
        plan.close()
        if expect == GENERAL:
            expect_fact_refused(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    else:
        fact, *_ = run_plan(w, FACT, from_orbitals=(c, e))
        assert np.abs(jac - fact).max() <= TOL
        nosym, *_ = run_plan(w, _lib.KERNEL_DMMA_NOSYM, from_orbitals=(c, e))
        assert not run_plan.last_symmetric
        assert np.abs(jac - nosym).max() <= 1e-12


def test_dmma_exact_block_symmetry_and_empty_perturbation():
    w = wl.sweep(128)
    # an empty perturbation (nnz_p = 0) in the middle of the list: its column must be exactly zero
    k = 57
    hs = np.concatenate([w.h1start[:k + 1], w.h1start[k:]]).astype(np.int32)
    w2 = wl.Workload(**{**w.__dict__})
    w2.h1start = hs
    jac, c, e, d, used = run_plan(w2, _lib.KERNEL_AUTO)
    assert used == DMMA
    assert np.all(jac[:, k] == 0.0)
    ref = oracle_jac(w, c, e)
    assert np.abs(np.delete(jac, k, axis=1) - ref).max() <= TOL
    blocks = jac.reshape(4, 4, w.norb // 4, -1, order="F")
    assert np.abs(blocks - blocks.transpose(1, 0, 2, 3)).max() == 0.0


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_from_given_orbitals_and_column_slices(kernel):
    w = wl.sweep(120)
    e, c = ro.eigh_sorted(w.h0)
    full, *_ = run_plan(w, KERNELS[kernel], from_orbitals=(c, e))
    ref = oracle_jac(w, c, e)
    assert np.abs(full - ref).max() <= TOL
    # column sharding: any slice equals the same columns of the full Jacobian -- bit for bit for the
    # general kernel; the DMMA kernel computes the full range in symmetric mode (mirrored patches), the
    # slices without it, so those agree to rounding only
    for lo, hi in ((0, 1), (7, 133), (150, 300), (299, 300), (40, 40)):
        part, *_ = run_plan(w, KERNELS[kernel], from_orbitals=(c, e), cols=(lo, hi))
        assert part.shape == (w.nresid, hi - lo)
        if kernel in ("general", "factorised") or hi == lo:
            assert np.array_equal(part, full[:, lo:hi])
        else:
            assert np.abs(part - full[:, lo:hi]).max() <= 1e-12
    if kernel == "auto":
        a, *_ = run_plan(w, _lib.KERNEL_DMMA_NOSYM, from_orbitals=(c, e))
        b, *_ = run_plan(w, _lib.KERNEL_DMMA_NOSYM, from_orbitals=(c, e), cols=(7, 133))
        assert np.array_equal(b, a[:, 7:133])


@pytest.mark.parametrize("nranks", [2, 3])
def test_fused_allgather_ranks_emulated_on_one_device(nranks):
    """The multi-GPU path (each rank executes its share of the block-pair list and stores every element
    into ALL ranks' Jacobian buffers) with the ranks emulated as plans on one device and plain device
    pointers in place of IPC-mapped peers: every buffer must end up holding the complete Jacobian."""
    w = wl.sweep(200)
    w.h1start = np.concatenate([w.h1start[:31], w.h1start[30:]]).astype(np.int32)   # + one empty perturbation
    e, c = ro.eigh_sorted(w.h0)
    ref, *_ = run_plan(w, _lib.KERNEL_AUTO, from_orbitals=(c, e))
    plans = []
    for r in range(nranks):
        p = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
        p.set_orbitals(c, e)
        assert len(p.export_jacobian()) == 64
        plans.append(p)
    ptrs = [p.device_ptr() for p in plans]
    for r, p in enumerate(plans):
        p.attach_peer_pointers(r, ptrs)
    for p in plans:
        p.compute()
    for p in plans:
        p.sync()
    assert sum(p.flops for p in plans) == pytest.approx(plans[0].flops * nranks, rel=0.02)
    for p in plans:
        assert p.symmetric_used
        assert np.array_equal(p.jacobian(), ref)
        p.close()


def _random_list_workload(seed, symmetric_mode):
    """Arbitrary CSR perturbations (lib/rhf_response.cc takes any list of unit elements): single
    unsymmetric elements E_rc, duplicated elements, 3- and 4-element perturbations, all inside 4-site groups."""
    rng = np.random.default_rng(seed)
    n, m = 72, 4
    w = wl.sweep(n)
    hs, hr, hc = [0], [], []
    groups = range(n // m) if symmetric_mode else list(range(0, n // m, 2))
    for g in groups:
        sites = np.arange(m * g, m * g + m)
        # a ring through the four sites: the group is one connected component
        hr += [int(s) for s in sites]; hc += [int(s) for s in np.roll(sites, -1)]; hs.append(len(hr))
        for _ in range(int(rng.integers(1, 7))):
            k = int(rng.integers(1, 5))
            chain = rng.choice(sites, size=k + 1)            # connected chain r0->r1->...: never split
            if rng.random() < 0.3:
                chain[1:] = chain[1]                         # duplicated / diagonal elements
            hr += list(map(int, chain[:-1])); hc += list(map(int, chain[1:])); hs.append(len(hr))
    w.h1start, w.h1row, w.h1col = (np.array(hs, dtype=np.int32), np.array(hr, dtype=np.int32),
                                   np.array(hc, dtype=np.int32))
    return w


@pytest.mark.parametrize("symmetric_mode", [True, False])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_arbitrary_element_lists(seed, symmetric_mode):
    w = _random_list_workload(seed, symmetric_mode)
    jac, c, e, d, used = run_plan(w, _lib.KERNEL_AUTO)
    assert used == DMMA and run_plan.last_symmetric == symmetric_mode
    ref = oracle_jac(w, c, e)
    assert np.abs(jac - ref).max() <= TOL
    gen, *_ = run_plan(w, _lib.KERNEL_GENERAL, from_orbitals=(c, e))
    assert np.abs(gen - ref).max() <= TOL
    # and through the compat symbol: the full response of the same list against the oracle's
    lib = _lib.load()
    n, p = w.norb, w.nterms
    h0 = np.array(w.h0.reshape(n * n), dtype=np.float64)
    out = np.ones(n * n * p)
    assert lib.rhf_response(n, p, w.num_pairs, _lib._ptr(w.h1start), _lib._ptr(w.h1row), _lib._ptr(w.h1col),
                            _lib._ptr(h0), _lib._ptr(out)) == 0
    resp, _ = ro.rhf_response(n, p, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0)
    assert np.abs(out.reshape(p, n, n) - resp).max() <= TOL


def test_stateless_entry_points_and_plan_reuse():
    """qcdmet_jacobian / qcdmet_jacobian_from_orbitals (host buffers in and out, the INTEGRATION.md binding):
    repeated calls with the same lists reuse the cached plan, a new list replaces it."""
    lib = _lib.load()

    def call(w, h0=None, orbitals=None, cols=None):
        lo, hi = cols if cols else (0, w.nterms)
        jac = np.empty((w.nresid, hi - lo), dtype=np.float64, order="F")
        dbl = np.empty(w.nresid)
        common = (w.norb, w.nterms, w.num_pairs, _lib._ptr(w.h1start), _lib._ptr(w.h1row), _lib._ptr(w.h1col))
        rows = (len(w.row_start) - 1, _lib._ptr(w.row_start), _lib._ptr(w.rows), _lib._ptr(w.row_idx))
        if orbitals is None:
            eig = np.empty(w.norb)
            h0 = np.ascontiguousarray(h0)
            rc = lib.qcdmet_jacobian(*common, _lib._ptr(h0), *rows, lo, hi, _lib._ptr(jac), _lib._ptr(dbl), _lib._ptr(eig))
        else:
            c = np.asfortranarray(orbitals[0])
            ee = np.ascontiguousarray(orbitals[1])
            rc = lib.qcdmet_jacobian_from_orbitals(*common, _lib._ptr(c), _lib._ptr(ee), *rows, lo, hi,
                                                   _lib._ptr(jac), _lib._ptr(dbl))
        assert rc == 0, lib.qcdmet_last_error()
        return jac

    w = wl.sweep(48)
    e, c = ro.eigh_sorted(w.h0)
    ref = oracle_jac(w, c, e)
    assert np.abs(call(w, h0=w.h0) - ref).max() <= TOL
    assert np.abs(call(w, orbitals=(c, e)) - ref).max() <= TOL                     # cached plan, other entry point
    assert np.abs(call(w, orbitals=(c, e), cols=(13, 77)) - ref[:, 13:77]).max() <= TOL
    h1 = w.h0 + 0.05 * np.diag(np.cos(np.arange(w.norb)))                          # same lists, new matrix
    e1, c1 = ro.eigh_sorted(h1)
    assert np.abs(call(w, h0=h1) - oracle_jac(w, c1, e1)).max() <= TOL
    w2 = wl.hubbard24(True, fit_imp_bath=True, l=6)                                # different lists: new plan
    e2, c2 = ro.eigh_sorted(w2.h0)
    assert np.abs(call(w2, h0=w2.h0) - oracle_jac(w2, c2, e2)).max() <= TOL
    assert np.abs(call(w, h0=w.h0) - ref).max() <= TOL
    assert lib.qcdmet_shutdown() == 0
    assert np.abs(call(w, h0=w.h0) - ref).max() <= TOL
    assert lib.qcdmet_shutdown() == 0


def test_lists_spanning_the_whole_matrix():
    """Perturbations with up to 8 elements anywhere in the matrix (+ an empty one): the fused DMMA kernel does
    not apply; factorised path and general kernel against the oracle."""
    rng = np.random.default_rng(7)
    w = wl.make_workload("rnd", 50, 19, [4, 3, 5, 4, 2], fit_imp_bath=False)
    hs, hr, hc = [0], [], []
    for _ in range(60):
        k = int(rng.integers(1, 9))
        hr += list(map(int, rng.integers(0, w.norb, k)))
        hc += list(map(int, rng.integers(0, w.norb, k)))
        hs.append(len(hr))
    hs.insert(30, hs[29])
    w.h1start, w.h1row, w.h1col = np.array(hs, np.int32), np.array(hr, np.int32), np.array(hc, np.int32)
    jac, c, e, d, used = run_plan(w, _lib.KERNEL_AUTO)
    assert used == FACT
    assert np.abs(jac - oracle_jac(w, c, e)).max() <= TOL
    assert np.all(jac[:, 29] == 0.0)
    gen, *_ = run_plan(w, GENERAL, from_orbitals=(c, e))
    assert np.abs(gen - oracle_jac(w, c, e)).max() <= TOL and np.all(gen[:, 29] == 0.0)


def test_degenerate_shapes():
    """Empty fragments, a fragment-free plan, zero perturbations, and a single perturbation."""
    w = wl.sweep(24)
    # fragments with zero rows interleaved
    rs = np.array([0, 4, 4, 8, 8, 8, 12], dtype=np.int32)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, rs, None, w.row_idx[:12])
    plan.set_h0(w.h0)
    plan.compute()
    jac, c, e = plan.jacobian(), plan.orbitals(), plan.eigenvalues()
    ref = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, rs, None, w.row_idx[:12])
    assert jac.shape == (48, w.nterms) and np.abs(jac - ref).max() <= TOL
    plan.close()
    # no perturbations: nothing to compute, empty Jacobian
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start[:1], w.h1row[:0], w.h1col[:0], w.row_start, None, w.row_idx)
    plan.set_h0(w.h0)
    plan.compute()
    assert plan.jacobian().shape == (w.nresid, 0)
    assert np.abs(plan.dblocks().reshape(-1, 16)[:, [0, 5, 10, 15]].sum() - 2 * w.num_pairs) < 1e-10   # trace of D
    plan.close()
    # one perturbation
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start[:2], w.h1row[:1], w.h1col[:1], w.row_start, None, w.row_idx)
    plan.set_h0(w.h0)
    plan.compute()
    ref = ro.jacobian_from_orbitals(plan.orbitals(), plan.eigenvalues(), w.num_pairs, w.h1start[:2], w.h1row[:1],
                                    w.h1col[:1], w.row_start, None, w.row_idx)
    assert np.abs(plan.jacobian() - ref).max() <= TOL
    plan.close()


def test_sn2_like_sampled_columns():
    w = wl.sn2_like()
    e, c = ro.eigh_sorted(w.h0)
    cols = (2000, 2040)
    part, _, _, _, used = run_plan(w, _lib.KERNEL_AUTO, from_orbitals=(c, e), cols=cols)
    assert used == FACT
    s = w.h1start
    ref = ro.jacobian_from_orbitals(c, e, w.num_pairs, s[cols[0]:cols[1] + 1] - s[cols[0]],
                                    w.h1row[s[cols[0]]:s[cols[1]]], w.h1col[s[cols[0]]:s[cols[1]]],
                                    w.row_start, w.rows, w.row_idx)
    assert np.abs(part - ref).max() <= TOL


def _column_sample_oracle(w, c, e, cols):
    s = w.h1start
    out = []
    for p in cols:
        out.append(ro.jacobian_from_orbitals(c, e, w.num_pairs, s[p:p + 2] - s[p], w.h1row[s[p]:s[p + 1]],
                                             w.h1col[s[p]:s[p + 1]], w.row_start, w.rows, w.row_idx)[:, 0])
    return np.stack(out, axis=1)


@pytest.mark.parametrize("shape", ["frag100_sliced", "frag102_bath_sliced", "frag89_whole_ragged"])
def test_factorised_large_fragments(shape):
    """Fragments beyond one contraction tile (> 96 rows: pairs of 8-block slices), ragged last blocks,
    projector rows: factorised path against the oracle on sampled columns, against the general kernel
    on all columns, exact block symmetry, column ranges bit-identical to the full run."""
    w = {"frag100_sliced": lambda: wl.make_workload("big", 150, 60, [100, 30, 9]),
         "frag102_bath_sliced": lambda: wl.make_workload("bb", 130, 55, [51, 3], fit_imp_bath=True),
         "frag89_whole_ragged": lambda: wl.make_workload("rg", 131, 57, [89, 1, 6, 2, 7, 13, 5]),
         }[shape]()
    e, c = ro.eigh_sorted(w.h0)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_orbitals(c, e)
    plan.compute()
    jac = plan.jacobian()
    assert plan.kernel_used == FACT and plan.launches == 2       # panel kernel + contraction
    cols = np.unique(np.linspace(0, w.nterms - 1, 24).astype(int))
    assert np.abs(jac[:, cols] - _column_sample_oracle(w, c, e, cols)).max() <= TOL
    off = 0
    for a, b in zip(w.row_start[:-1], w.row_start[1:]):
        r = int(b - a)
        blk = jac[off:off + r * r].reshape(r, r, -1, order="F")
        assert np.array_equal(blk, blk.transpose(1, 0, 2))
        off += r * r
    lo, hi = w.nterms // 3, w.nterms // 3 + 17
    plan.set_columns(lo, hi)
    plan.compute()
    assert plan.launches == 2
    assert np.array_equal(plan.jacobian(), jac[:, lo:hi])
    plan.set_columns(0, w.nterms)
    plan.set_h0(w.h0 + 0.01 * np.eye(w.norb))
    plan.compute()
    assert plan.launches == 2
    assert np.abs(plan.jacobian() - jac).max() <= 1e-9          # a constant shift does not change the response
    plan.set_kernel(GENERAL)
    plan.set_orbitals(c, e)
    plan.compute()
    assert plan.kernel_used == GENERAL
    assert np.abs(plan.jacobian() - jac).max() <= TOL
    plan.close()


def test_compat_full_response_sliced():
    """rhf_response at a size where the single all-sites fragment (N = 100: 13 blocks) is cut into slice pairs."""
    w = wl.sweep(100, nterms=23)
    lib = _lib.load()
    n, p = w.norb, w.nterms
    h0 = np.array(w.h0.reshape(n * n), dtype=np.float64)
    out = np.ones(n * n * p)
    assert lib.rhf_response(n, p, w.num_pairs, _lib._ptr(w.h1start), _lib._ptr(w.h1row), _lib._ptr(w.h1col),
                            _lib._ptr(h0), _lib._ptr(out)) == 0, lib.qcdmet_last_error()
    resp, one_rdm = ro.rhf_response(n, p, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0)
    out = out.reshape(p, n, n)
    assert np.abs(out - resp).max() <= TOL
    assert np.array_equal(out, out.transpose(0, 2, 1))
    assert np.abs(h0.reshape(n, n) - one_rdm).max() <= TOL


@pytest.mark.skipif(os.environ.get("QCDMET_TEST_EXPERIMENTAL") != "1",
                    reason="opt-in warp-specialised contraction kernel: set QCDMET_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("shape", ["sn2", "polyyne", "frag100", "ti_m5", "random_lists"])
def test_experimental_ws_contraction_is_bit_identical(shape, monkeypatch):
    """QCDMET_B200_CONTRACT_WS=1 (staging warps + DMMA warps) must reproduce the default
    contraction kernel bit for bit: same schedule, same operation order per accumulator."""
    w = {"sn2": wl.sn2_like, "polyyne": wl.polyyne_like,
         "frag100": lambda: wl.make_workload("big", 150, 60, [100, 30, 9], nterms=700),
         "ti_m5": lambda: wl.make_workload("ti5", 40, 18, [5] * 8, trans_inv=True),
         "random_lists": lambda: _random_list_workload(1, True)}[shape]()
    if shape == "sn2":
        w = wl.make_workload("sn2s", 321, 58, [52] + [24] * 10 + [29], nterms=600, seed=20240 + 321)
    e, c = ro.eigh_sorted(w.h0)
    monkeypatch.delenv("QCDMET_B200_CONTRACT_WS", raising=False)
    ref, *_ = run_plan(w, FACT, from_orbitals=(c, e))
    monkeypatch.setenv("QCDMET_B200_CONTRACT_WS", "1")
    got, *_ = run_plan(w, FACT, from_orbitals=(c, e))
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("shape", ["sn2", "polyyne", "frag100", "small_no", "column_shard"])
def test_resident_contraction_matches_chunked(shape, monkeypatch):
    """Occupied spaces of <= 64 orbitals with <= 2 elements per perturbation run contract_resident_kernel (PO of
    the tile and the panel of the first element's row site stay in shared memory, the second panel is streamed
    through registers); QCDMET_B200_CONTRACT_RESIDENT=0 selects the chunked kernel.  Same schedule and stores; the
    resident kernel keeps the two products of a block pair in separate accumulators (added at the end of a column),
    so the results agree to rounding, blocks stay exactly symmetric and runs are bit-reproducible."""
    w = {"sn2": lambda: wl.make_workload("sn2s", 321, 58, [52] + [24] * 10 + [29], nterms=900, seed=20240 + 321),
         "polyyne": wl.polyyne_like,
         "frag100": lambda: wl.make_workload("big", 150, 60, [100, 30, 9], nterms=700),
         "small_no": lambda: wl.make_workload("few", 90, 11, [17, 40, 9, 24], nterms=333),
         "column_shard": lambda: wl.make_workload("sn2c", 200, 40, [52, 24, 29, 24, 24, 47], nterms=500)}[shape]()
    e, c = ro.eigh_sorted(w.h0)
    cols = (123, 407) if shape == "column_shard" else None
    monkeypatch.setenv("QCDMET_B200_CONTRACT_RESIDENT", "0")
    ref, *_ = run_plan(w, FACT, from_orbitals=(c, e), cols=cols)
    monkeypatch.delenv("QCDMET_B200_CONTRACT_RESIDENT")
    got, *_ = run_plan(w, FACT, from_orbitals=(c, e), cols=cols)
    again, *_ = run_plan(w, FACT, from_orbitals=(c, e), cols=cols)
    assert np.array_equal(got, again)                       # run-to-run bit-reproducible
    assert np.abs(got - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    # both tile sizes of the resident kernel (64 row slots x 2 CTAs per SM / 128 slots) run the same arithmetic per
    # accumulator: bit-identical to each other
    for slots in ("64", "128"):
        monkeypatch.setenv("QCDMET_B200_CONTRACT_SLOTS", slots)
        other, *_ = run_plan(w, FACT, from_orbitals=(c, e), cols=cols)
        assert np.array_equal(other, got), slots
    monkeypatch.delenv("QCDMET_B200_CONTRACT_SLOTS")
    lo, hi = cols if cols else (0, w.nterms)
    sample = np.unique(np.linspace(lo, hi - 1, 12).astype(int))
    lens = w.h1start[sample + 1] - w.h1start[sample]
    hs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    take = np.concatenate([np.arange(w.h1start[p], w.h1start[p + 1]) for p in sample])
    orc = ro.jacobian_from_orbitals(c, e, w.num_pairs, hs, w.h1row[take], w.h1col[take], w.row_start, w.rows, w.row_idx)
    assert np.abs(got[:, sample - lo] - orc).max() <= TOL


@pytest.mark.parametrize("shape", ["sweep400", "sweep576", "hubbard24_bath", "sweep1000"])
def test_tile_split_balancing_is_bit_identical(shape, monkeypatch):
    """When the block-pair list is not a multiple of the SM count the fused kernel balances the SMs at column-tile
    granularity (SPLIT instantiation: an item cut by a range boundary is finished by the CTA holding its first tiles,
    which adds the other CTA's per-tile contributions in tile order).  Same association as the unsplit item: bit-identical
    Jacobians, whatever the split."""
    w = {"sweep400": lambda: wl.sweep(400), "sweep576": lambda: wl.sweep(576),
         "hubbard24_bath": lambda: wl.hubbard24(fit_imp_bath=True, l=24),       # 8-row fragments: one running-sum buffer
         "sweep1000": lambda: wl.sweep(1000)}[shape]()
    e, c = ro.eigh_sorted(w.h0)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("QCDMET_B200_SPLIT", flag)
        plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
        plan.set_orbitals(c, e)
        plan.compute()
        out[flag] = (plan.jacobian(), plan.split_used, plan.kernel_used)
        again = None
        if flag == "1":
            plan.compute()
            again = plan.jacobian()
            assert np.array_equal(again, out[flag][0])          # bit-reproducible run to run
        plan.close()
    assert out["0"][2] == _lib.KERNEL_DMMA and not out["0"][1]
    assert out["1"][1], "the split instantiation did not run"
    assert np.array_equal(out["0"][0], out["1"][0])
    sample = np.unique(np.linspace(0, w.nterms - 1, 6).astype(int))
    lens = w.h1start[sample + 1] - w.h1start[sample]
    hs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    take = np.concatenate([np.arange(w.h1start[p], w.h1start[p + 1]) for p in sample])
    orc = ro.jacobian_from_orbitals(c, e, w.num_pairs, hs, w.h1row[take], w.h1col[take], w.row_start, w.rows, w.row_idx)
    assert np.abs(out["1"][0][:, sample] - orc).max() <= TOL


def test_pageable_and_page_locked_host_copies_agree():
    """A fresh (pageable) numpy destination of more than 128 MB takes the pipelined staged copy (two page-locked 32 MB
    buffers + host threads), a page-locked destination the direct copy: same bytes."""
    w = wl.sweep(1600)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_h0(w.h0)
    plan.compute()
    assert w.nresid * w.nterms * 8 > 4 * (32 << 20)
    fresh = plan.jacobian()                                   # new pageable array
    pinned = plan.jacobian(pinned=True)                       # the plan's page-locked buffer
    assert fresh.shape == pinned.shape == (w.nresid, w.nterms)
    assert np.array_equal(fresh, pinned)
    part = plan.jacobian_columns(1234, 3999)                  # odd offsets, still above the threshold
    assert np.array_equal(part, pinned[:, 1234:3999])
    plan.close()


def test_invariants_large():
    """Size-independent properties at a size the CPU oracle does not reach in seconds."""
    w = wl.sweep(1000, nterms=600)
    a, c, e, d, used_a = run_plan(w, _lib.KERNEL_AUTO)
    b, *_ = run_plan(w, _lib.KERNEL_GENERAL, from_orbitals=(c, e))
    assert np.abs(a - b).max() <= TOL           # two independent kernels agree
    q = w.nresid
    jac = a
    # each 4x4 block is symmetric and the full dD_p is traceless: with fragments tiling all sites,
    # the sum of the diagonal entries over all blocks vanishes
    blocks = jac.reshape(4, 4, w.norb // 4, w.nterms, order="F")
    assert np.abs(blocks - blocks.transpose(1, 0, 2, 3)).max() == 0.0
    tr = np.einsum("aafp->p", blocks)
    assert np.abs(tr).max() < 1e-11
    # Gram symmetry of the static response function: rows (f;a,b) and columns p=(g;c<=d) index the
    # same operator pairs; the column operator is E_cd+E_dc = (2-delta_cd) x the row operator.
    pair_rows, scale = [], []
    for f in range(w.norb // 4):
        for r in range(4):
            for cc in range(r, 4):
                pair_rows.append(16 * f + r + 4 * cc)
                scale.append(1.0 if r == cc else 0.5)
    sq = jac[np.array(pair_rows[:w.nterms]), :] * np.array(scale[:w.nterms])[None, :]
    assert np.abs(sq - sq.T).max() < 1e-11


def test_north_star_size_properties():
    """The full target: Norb = 4000, 10^4 perturbations, Q = 16000 (1.28 GB Jacobian).  The oracle cannot
    produce it in seconds, so: size-independent properties on the whole matrix, the second (general) kernel
    on a slice of columns, and the oracle on two columns."""
    w = wl.sweep(4000)
    assert (w.nterms, w.nresid) == (10000, 16000)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_h0(w.h0)
    plan.compute()
    jac = plan.jacobian()
    assert plan.kernel_used == DMMA and plan.symmetric_used
    c, e = plan.orbitals(), plan.eigenvalues()
    blocks = jac.reshape(4, 4, 1000, 10000, order="F")
    for lo in range(0, 10000, 2500):                 # exact block symmetry, chunked to bound host memory
        b = blocks[:, :, :, lo:lo + 2500]
        assert np.array_equal(b, b.transpose(1, 0, 2, 3))
    tr = np.einsum("aafp->p", blocks)
    assert np.abs(tr).max() < 1e-10                  # dD_p is traceless (fragments tile every site)
    # Gram symmetry of the response function on a random sample of (row pair, column pair) entries
    rng = np.random.default_rng(0)
    pair_row = np.array([16 * f + r + 4 * cc for f in range(1000) for r in range(4) for cc in range(r, 4)])
    scale = np.array([1.0 if r == cc else 0.5 for f in range(1000) for r in range(4) for cc in range(r, 4)])
    i, j = rng.integers(0, 10000, 4000), rng.integers(0, 10000, 4000)
    assert np.abs(jac[pair_row[i], j] * scale[j] - jac[pair_row[j], i] * scale[i]).max() < 1e-11
    # sum rule: D stays idempotent, D dD + dD D = 2 dD  =>  for the diagonal blocks sum_b ... (checked on D itself)
    d = plan.dblocks().reshape(4, 4, 1000, order="F")
    assert abs(np.einsum("aaf->", d) - 2 * w.num_pairs) < 1e-9
    # the independent general kernel on 24 columns spread over the matrix
    plan.set_kernel(_lib.KERNEL_GENERAL)
    for lo in (0, 4990, 9992):
        plan.set_columns(lo, lo + 8)
        plan.compute()
        part = plan.jacobian()
        assert np.abs(part - jac[:, lo:lo + 8]).max() <= TOL
    plan.close()
    # the oracle on two columns (a diagonal and a pair perturbation of a fragment in the middle)
    s = w.h1start
    for p in (5000, 5001):
        ref = ro.jacobian_from_orbitals(c, e, w.num_pairs, s[p:p + 2] - s[p], w.h1row[s[p]:s[p + 1]],
                                        w.h1col[s[p]:s[p + 1]], w.row_start, None, w.row_idx)
        assert np.abs(ref[:, 0] - jac[:, p]).max() <= TOL


def test_bitwise_reproducible():
    """No atomics, fixed summation order: repeated runs (and a fresh plan) give identical bits."""
    for w in (wl.sweep(200), wl.hubbard24(True, l=8), wl.polyyne_like()):
        e, c = ro.eigh_sorted(w.h0)
        a, *_ = run_plan(w, _lib.KERNEL_AUTO, from_orbitals=(c, e))
        plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
        plan.set_orbitals(c, e)
        for _ in range(3):
            plan.compute()
            assert np.array_equal(plan.jacobian(), a)
        plan.close()


def test_linearity_in_perturbation():
    """dD is linear in H1: a TI perturbation equals the sum of its per-cell perturbations."""
    l = 8
    wt = wl.hubbard24(True, l=l)
    wn = wl.hubbard24(False, l=l)
    e, c = ro.eigh_sorted(wt.h0)
    # rows: fragment 0 only, in both
    wn_rows = wl.Workload(**{**wn.__dict__})
    wn_rows.row_start = wt.row_start
    wn_rows.row_idx = wt.row_idx
    jt, *_ = run_plan(wt, _lib.KERNEL_AUTO, from_orbitals=(c, e))
    jn, *_ = run_plan(wn_rows, _lib.KERNEL_AUTO, from_orbitals=(c, e))
    ncell = l * l // 4
    summed = jn.reshape(16, ncell, 10).sum(axis=1)
    assert np.abs(jt - summed).max() <= TOL


def test_state_errors_of_the_multi_gpu_and_lm_entry_points():
    w = wl.polyyne_like()                          # general kernel only
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_h0(w.h0)
    with pytest.raises(_lib.QcdmetError) as ei:
        plan.lm_step(np.zeros(w.nresid), 1.0)      # no Jacobian yet
    assert ei.value.code == -5
    plan.export_jacobian()
    with pytest.raises(_lib.QcdmetError) as ei:
        plan.attach_peer_pointers(0, [plan.device_ptr(), plan.device_ptr()])    # needs the DMMA kernel
    assert ei.value.code == -5
    plan.close()
    w = wl.sweep(32)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_h0(w.h0)
    plan.set_columns(0, 10)
    with pytest.raises(_lib.QcdmetError) as ei:
        plan.export_jacobian()                     # peer mode works on the full column range
    assert ei.value.code == -5
    plan.set_columns(0, w.nterms)
    plan.export_jacobian()
    with pytest.raises(_lib.QcdmetError) as ei:
        plan.attach_peer_pointers(3, [plan.device_ptr()] * 2)
    assert ei.value.code == -1
    plan.close()


def test_gap_error_is_reported():
    w = wl.sweep(16)
    h0 = np.zeros((16, 16))    # fully degenerate: no HOMO-LUMO gap
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    with pytest.raises(_lib.QcdmetError) as ei:
        plan.set_h0(h0)
    assert ei.value.code == -6
    with pytest.raises(_lib.QcdmetError) as ei:
        plan.compute()
    assert ei.value.code == -5
    plan.close()


class Ints:
    TI_OK = True
    ERIinMEM = False

    def __init__(self, fock, nelec):
        self._f, self.Norbs, self.Nelec = np.asarray(fock), fock.shape[0], nelec

    def loc_rhf_fock(self):
        return self._f

    def loc_oei(self):
        return self._f


@pytest.mark.parametrize("case", ["imp_only", "imp_bath", "ti_bath"])
def test_dmet_class_jacobian_and_gradient(case):
    """dmet.rdm_differences_derivative (reference name/shape) against the oracle's restatement of
    src/dmet.py:477-525, at a non-zero umat, plus the verify_gradient check of src/dmet.py:527-539."""
    w = {"imp_only": lambda: wl.make_workload("m", 40, 17, [4, 6, 3, 5]),
         "imp_bath": lambda: wl.make_workload("b", 36, 15, [3, 4, 2], fit_imp_bath=True),
         "ti_bath": lambda: wl.hubbard24(True, fit_imp_bath=True, l=6)}[case]()
    d = dmet_mod.dmet(Ints(w.h0, 2 * w.num_pairs), w.imp_clust, w.trans_inv, fitImpBath=w.fit_imp_bath)
    if w.fit_imp_bath:
        d.dmetOrbs = w.dmet_orbs
    rng = np.random.default_rng(5)
    nfit = len(d.imp_size)
    sizes = [b.shape[1] for b in w.dmet_orbs] if w.fit_imp_bath else [2 * int(m) for m in d.imp_size]
    d.imp_1RDM = [(lambda a: a + a.T)(rng.standard_normal((s, s))) for s in sizes[:nfit]]
    x = rng.standard_normal(d.helper.Nterms) * 0.02
    jac = d.rdm_differences_derivative(x)
    ref = ro.rdm_differences_derivative(w.h0, x, w.num_pairs, w.imp_clust, w.trans_inv, w.fit_imp_bath, w.dmet_orbs)
    assert jac.shape == ref.shape
    assert np.abs(jac - ref).max() <= TOL
    diff, norm = d.verify_gradient(x)
    assert diff < 1e-4 * max(norm, 1.0)
    g_dev = d.costfunction_derivative(x)                       # 2 J^T r on the device
    assert np.abs(g_dev - 2.0 * jac.T @ d.rdm_differences(x)).max() <= 1e-11 * max(1.0, np.abs(g_dev).max())
    assert d.time_grad > 0.0


@pytest.mark.parametrize("case", ["imp_only", "imp_bath", "ti_bath"])
def test_gpu_residual_matches_host_and_shares_eigensolve(case):
    """dmet.rdm_differences from the device eigensolve (f-2) against the reference's host route and the
    oracle; one syevd serves the residual and the Jacobian at the same umat."""
    w = {"imp_only": lambda: wl.make_workload("m", 40, 17, [4, 6, 3, 5]),
         "imp_bath": lambda: wl.make_workload("b", 36, 15, [3, 4, 2], fit_imp_bath=True),
         "ti_bath": lambda: wl.hubbard24(True, fit_imp_bath=True, l=6)}[case]()
    d = dmet_mod.dmet(Ints(w.h0, 2 * w.num_pairs), w.imp_clust, w.trans_inv, fitImpBath=w.fit_imp_bath)
    if w.fit_imp_bath:
        d.dmetOrbs = w.dmet_orbs
    rng = np.random.default_rng(11)
    nfit = len(d.imp_size)
    sizes = [b.shape[1] for b in w.dmet_orbs] if w.fit_imp_bath else [2 * int(m) for m in d.imp_size]
    d.imp_1RDM = [(lambda a: a + a.T)(rng.standard_normal((s, s))) for s in sizes[:nfit]]
    x = rng.standard_normal(d.helper.Nterms) * 0.02
    r_gpu = d.rdm_differences(x)
    assert d.helper.num_eigensolves == 1
    jac = d.rdm_differences_derivative(x)
    assert d.helper.num_eigensolves == 1                 # same umat: the eigensolve is reused
    r_host = d.rdm_differences_host(x)
    r_ref = ro.rdm_differences(w.h0, x, w.num_pairs, w.imp_clust, w.trans_inv, w.fit_imp_bath, d.imp_1RDM, w.dmet_orbs)
    assert np.abs(r_gpu - r_ref).max() <= TOL and np.abs(r_host - r_ref).max() <= TOL
    d.rdm_differences(x + 1e-3)
    assert d.helper.num_eigensolves == 2
    assert jac.shape == (r_gpu.size, d.helper.Nterms)


def test_helper_full_response_api():
    w = wl.make_workload("m", 24, 10, [4, 4, 4])
    d = dmet_mod.dmet(Ints(w.h0, 20), w.imp_clust, False, fitImpBath=False)
    resp = d.helper.construct1RDM_response(False, np.zeros((24, 24)), None)
    ref, _ = ro.rhf_response(24, w.nterms, 10, w.h1start, w.h1row, w.h1col, w.h0)
    assert resp.shape == (w.nterms, 24, 24)
    assert np.abs(resp - ref).max() <= TOL
    resp2 = d.helper.construct1RDM_loc_response(False, np.zeros((24, 24)), None)
    assert np.array_equal(resp, resp2)
