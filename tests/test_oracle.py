"""CPU tests: the oracle against the reference's own outputs (golden fixtures generated from the
compiled, unmodified rhf_response.cc; and that library itself when oracle/_ref is present)."""
import numpy as np
import pytest

from oracle import ref_loader, response_oracle as ro
from qcdmet_b200 import workloads as wl
from util import GOLDEN, golden_ids, load_golden

TOL = 1e-10   # BASELINE.json north_star: 1e-10 absolute on response elements


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_oracle_matches_golden_response(path):
    g = load_golden(path)
    resp, one_rdm = ro.rhf_response(g["norb"], len(g["h1start"]) - 1, g["num_pairs"], g["h1start"],
                                    g["h1row"], g["h1col"], g["h0"])
    assert np.abs(resp - g["resp"]).max() <= TOL
    assert np.abs(one_rdm - g["one_rdm"]).max() <= TOL


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_oracle_jacobian_matches_golden(path):
    g = load_golden(path)
    e, c = ro.eigh_sorted(g["h0"])
    jac = ro.jacobian_from_orbitals(c, e, g["num_pairs"], g["h1start"], g["h1row"], g["h1col"],
                                    g["row_start"], g["rows"], g["row_idx"])
    assert jac.shape == g["jac"].shape
    assert np.abs(jac - g["jac"]).max() <= TOL


@pytest.mark.skipif(not ref_loader.available("mkl"), reason="oracle/_ref not built")
@pytest.mark.parametrize("flavour", ["mkl", "openblas"])
def test_oracle_matches_compiled_reference(flavour):
    if not ref_loader.available(flavour):
        pytest.skip("flavour not built")
    for w in (wl.sweep(48), wl.make_workload("odd", 37, 13, [6, 1, 5, 2]), wl.hubbard24(True, l=8)):
        resp_ref, d_ref = ref_loader.rhf_response(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0,
                                                  flavour)
        resp, d = ro.rhf_response(w.norb, w.nterms, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0)
        assert np.abs(resp - resp_ref).max() <= TOL
        assert np.abs(d - d_ref).max() <= TOL


@pytest.mark.skipif(not ref_loader.available("mkl"), reason="oracle/_ref not built")
def test_oracle_matches_compiled_reference_on_arbitrary_lists():
    """lib/rhf_response.cc accepts any list of unit elements: unsymmetric single elements, duplicates,
    long lists spanning the whole matrix.  The restatement must agree on all of them."""
    rng = np.random.default_rng(7)
    n, no = 26, 9
    w = wl.make_workload("rnd", n, no, [4, 4, 4])
    hs, hr, hc = [0], [], []
    for _ in range(40):
        k = int(rng.integers(1, 9))
        hr += list(map(int, rng.integers(0, n, k)))
        hc += list(map(int, rng.integers(0, n, k)))
        hs.append(len(hr))
    hs.append(len(hr))                                  # an empty perturbation at the end
    hs, hr, hc = np.array(hs, np.int32), np.array(hr, np.int32), np.array(hc, np.int32)
    resp_ref, d_ref = ref_loader.rhf_response(n, no, hs, hr, hc, w.h0, "mkl")
    resp, d = ro.rhf_response(n, len(hs) - 1, no, hs, hr, hc, w.h0)
    assert np.abs(resp - resp_ref).max() <= TOL and np.abs(d - d_ref).max() <= TOL
    assert np.all(resp[-1] == 0.0) and np.all(resp_ref[-1] == 0.0)
    e, c = ro.eigh_sorted(w.h0)
    jac = ro.jacobian_from_orbitals(c, e, no, hs, hr, hc, w.row_start, None, w.row_idx)
    imp_size = ro.make_imp_size(w.imp_clust, False)
    assert np.abs(jac - ro.jacobian_from_response(resp_ref, w.imp_clust, imp_size, False)).max() <= TOL


@pytest.mark.skipif(not ref_loader.available("mkl"), reason="oracle/_ref not built")
def test_compiled_reference_nterms_zero():
    w = wl.sweep(16, nterms=0)
    resp, d = ref_loader.rhf_response(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0)
    assert resp.shape == (0, 16, 16)
    assert abs(np.trace(d) - 2 * w.num_pairs) < 1e-12


def test_oracle_finite_difference():
    """dD_p is the derivative of 2 C_o C_o^T w.r.t. the strength of H1_p (central differences)."""
    w = wl.make_workload("fd", 14, 6, [3, 4, 2])
    resp, _ = ro.rhf_response(w.norb, w.nterms, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0)
    eps = 1e-6
    for p in (0, 5, w.nterms - 1):
        h1 = np.zeros((w.norb, w.norb))
        h1[w.h1row[w.h1start[p]:w.h1start[p + 1]], w.h1col[w.h1start[p]:w.h1start[p + 1]]] = 1.0
        fd = (ro.construct1rdm_base(w.h0 + eps * h1, w.num_pairs)
              - ro.construct1rdm_base(w.h0 - eps * h1, w.num_pairs)) / (2 * eps)
        assert np.abs(fd - resp[p]).max() < 1e-7


def test_oracle_invariants():
    w = wl.sweep(32)
    resp, d = ro.rhf_response(w.norb, w.nterms, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0)
    assert abs(np.trace(d) - 2 * w.num_pairs) < 1e-12
    for p in range(0, w.nterms, 7):
        dd = resp[p]
        assert np.abs(dd - dd.T).max() == 0.0                      # exactly symmetric
        assert abs(np.trace(dd)) < 1e-12                           # particle number conserved
        assert np.abs(d @ dd + dd @ d - 2 * dd).max() < 1e-12      # D^2 = 2D differentiated


def test_parameter_ordering_and_masks():
    for w in (wl.sweep(16), wl.make_workload("m", 20, 8, [3, 5, 2]), wl.hydrogen_like()):
        imp_size = ro.make_imp_size(w.imp_clust, w.trans_inv)
        s, r, c = ro.makelist_h1_sparse(w.norb, imp_size, w.trans_inv)
        s2, r2, c2 = ro.convert_h1_sparse(ro.makelist_h1(w.norb, imp_size, w.trans_inv))
        assert (s == s2).all() and (r == r2).all() and (c == c2).all()
        assert (s == w.h1start).all() and (r == w.h1row).all() and (c == w.h1col).all()
        mask = ro.make_mask(w.norb, imp_size)
        assert mask.sum() == len(s) - 1
        x = np.arange(1, len(s)) * 0.5
        u = ro.flat2square(x, w.norb, mask, imp_size, w.trans_inv)
        assert np.abs(u - u.T).max() == 0
        assert np.allclose(ro.square2flat(u, mask), x)
        # flat2square == sum_p x_p H1_p   (the commented-out check at src/dmet.py:571-575)
        dense = ro.makelist_h1(w.norb, imp_size, w.trans_inv)
        assert np.allclose(u, sum(xp * h for xp, h in zip(x, dense)))


def test_jacobian_is_derivative_of_residual():
    w = wl.make_workload("res", 18, 7, [4, 3, 5])
    rng = np.random.default_rng(3)
    imp_1rdm = [rng.standard_normal((2 * m, 2 * m)) for m in (4, 3, 5)]
    x = rng.standard_normal(w.nterms) * 0.05
    fock = w.h0
    jac = ro.rdm_differences_derivative(fock, x, w.num_pairs, w.imp_clust, False, False)
    eps = 1e-6
    for p in (0, 9, w.nterms - 1):
        dx = np.zeros_like(x)
        dx[p] = eps
        fd = (ro.rdm_differences(fock, x + dx, w.num_pairs, w.imp_clust, False, False, imp_1rdm)
              - ro.rdm_differences(fock, x - dx, w.num_pairs, w.imp_clust, False, False, imp_1rdm)) / (2 * eps)
        assert np.abs(fd - jac[:, p]).max() < 1e-7
