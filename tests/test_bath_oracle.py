"""The bath / embedding checker (oracle/bath_oracle.py) against the properties a Schmidt decomposition of
a closed-shell determinant must have, and the product's host mirror (qcdmethelper.constructbath,
localintegrals_hubbard.dmet_*) against the checker.  No GPU."""
import numpy as np
import pytest

from oracle import bath_oracle as bo
from qcdmet_b200 import localintegrals_hubbard as lh, qcdmethelper as qh, workloads as wl


def determinant(n, num_pairs, seed):
    h0, _, _ = wl.spectrum_matrix(n, num_pairs, seed)
    e, c = np.linalg.eigh(h0)
    return h0, 2.0 * c[:, :num_pairs] @ c[:, :num_pairs].T


def flags_of(n, idx):
    f = np.zeros(n, dtype=int)
    f[np.asarray(idx)] = 1
    return f


CASES = [
    (24, 10, [1, 2, 3, 6]),
    (40, 17, [0, 5, 6, 7, 30]),
    (30, 4, [2, 3, 4, 5, 6, 7, 8]),              # more impurity orbitals than electron pairs: only 4 entangled orbitals
    (30, 27, [2, 3, 4, 5, 6, 7, 8]),             # ... than virtual orbitals: only 3
    (16, 8, list(range(0, 16, 2))),              # half of the lattice
]


@pytest.mark.parametrize("n,num_pairs,imp", CASES)
def test_oracle_bath_is_a_schmidt_decomposition(n, num_pairs, imp):
    _, one = determinant(n, num_pairs, 3 + n)
    m = len(imp)
    nb, loc2dmet, core = bo.construct_bath(one, flags_of(n, imp), m)
    assert nb == min(m, num_pairs, n - num_pairs)
    assert np.abs(loc2dmet.T @ loc2dmet - np.eye(n)).max() < 1e-12
    assert np.abs(loc2dmet[imp, :m] - np.eye(m)).max() == 0.0           # impurity unit vectors first
    assert np.abs(loc2dmet[imp, m:]).max(initial=0.0) == 0.0                       # nothing else lives on the impurity
    rot = loc2dmet.T @ one @ loc2dmet
    act = m + nb
    # the determinant factorises: no coupling between the active space and the pure environment, whose orbitals are
    # doubly occupied or empty, sorted by decreasing occupation; the active block is an idempotent closed-shell 1-RDM
    assert np.abs(rot[:act, act:]).max(initial=0.0) < 1e-10
    assert np.abs(np.diag(rot)[act:] - core[act:]).max(initial=0.0) < 1e-10
    assert np.all(np.diff(core[act:]) <= 1e-12)
    assert np.abs(np.minimum(core[act:], 2.0 - core[act:])).max(initial=0.0) < 1e-10
    assert np.all(core[:act] == 0.0)
    pa = rot[:act, :act] / 2.0
    assert np.abs(pa @ pa - pa).max() < 1e-10
    nelec_active = np.trace(rot[:act, :act])
    assert abs(nelec_active - round(nelec_active)) < 1e-9 and int(round(nelec_active)) % 2 == 0
    # bath occupations come in the reference's order: closest to 1 first
    bocc = np.diag(rot)[m:act]
    assert np.all(np.diff(np.minimum(bocc, 2.0 - bocc)) <= 1e-10)
    dm, occ = bo.core_density(loc2dmet, core)
    assert abs(np.trace(dm) + nelec_active - 2 * num_pairs) < 1e-9
    assert set(np.unique(occ)) <= {0.0, 2.0}


@pytest.mark.parametrize("n,num_pairs,imp", CASES)
def test_host_constructbath_matches_oracle(n, num_pairs, imp):
    _, one = determinant(n, num_pairs, 3 + n)
    m = len(imp)
    helper = qh.qcdmethelper.__new__(qh.qcdmethelper)                   # constructbath uses no state
    for request in (m, max(1, m // 2)):
        nb, loc2dmet, core = qh.qcdmethelper.constructbath(helper, one, flags_of(n, imp), request)
        nb_o, loc_o, core_o = bo.construct_bath(one, flags_of(n, imp), request)
        assert nb == nb_o
        assert np.abs(core - core_o).max() < 1e-12
        act = m + nb
        got, ref = loc2dmet[:, m:act], loc_o[:, m:act]
        dots = np.sum(got * ref, axis=0)
        assert np.abs(np.abs(dots) - 1.0).max() < 1e-10                  # same bath orbitals up to a sign
        assert np.abs(got * np.sign(dots) - ref).max() < 1e-9
        # the pure environment is degenerate (occupations 0 and 2): compare the two subspaces
        for value in (0.0, 2.0):
            a = loc2dmet[:, act:][:, np.abs(core[act:] - value) < 1e-8]
            b = loc_o[:, act:][:, np.abs(core_o[act:] - value) < 1e-8]
            assert a.shape == b.shape
            assert np.abs(a @ a.T - b @ b.T).max(initial=0.0) < 1e-9


def test_threshold_drops_unentangled_orbitals():
    n, num_pairs = 20, 8
    h0, _, _ = wl.spectrum_matrix(n, num_pairs, 9)
    h0[4, :] = 0.0
    h0[:, 4] = 0.0
    h0[4, 4] = -7.0                                                      # isolated, doubly occupied site
    e, c = np.linalg.eigh(h0)
    one = 2.0 * c[:, :num_pairs] @ c[:, :num_pairs].T
    imp = [3, 4, 5]
    nb, loc2dmet, core = bo.construct_bath(one, flags_of(n, imp), 3)
    assert nb == 2                                                       # src/qcdmethelper.py:140-143
    helper = qh.qcdmethelper.__new__(qh.qcdmethelper)
    nb_h, _, core_h = qh.qcdmethelper.constructbath(helper, one, flags_of(n, imp), 3)
    assert nb_h == 2 and np.abs(core - core_h).max() < 1e-12


def test_hubbard_embedding_matches_oracle():
    n, nelec, u = 12, 12, 4.0
    hop = lh.hubbard_chain(n)
    ints = lh.localintegrals_hubbard(hop, u, nelec)
    one = ints.origDMloc
    imp = [4, 5]
    nb, loc2dmet, core = bo.construct_bath(one, flags_of(n, imp), 2)
    act = 2 + nb
    core_dm, _ = bo.core_density(loc2dmet, core)
    assert np.abs(ints.dmet_oei(loc2dmet, act) - bo.embedded_oei(hop, loc2dmet, act)).max() < 1e-13
    assert np.abs(ints.dmet_fock(loc2dmet, act, core_dm) - bo.embedded_fock(hop, u, loc2dmet, act, core_dm)).max() < 1e-13
    assert np.abs(ints.dmet_tei(loc2dmet, act) - bo.embedded_tei(u, loc2dmet, act)).max() < 1e-13
    assert np.abs(ints.loc_rhf_fock_bis(one) - bo.fock_from_core(hop, u, one)).max() == 0.0
    # half filling on a bipartite lattice: every site carries one electron, the mean-field term is U/2 on the diagonal
    assert np.abs(np.diag(bo.fock_from_core(hop, u, one)) - u / 2).max() < 1e-12
