"""DMET self-consistency loop: CPU sanity checks of the restated host orchestration, and (GPU) parity of
the converged energy / umat between the oracle-driven and the GPU-driven loop: 1e-8 Hartree on energies
(BASELINE.json north_star)."""
import numpy as np
import pytest

from qcdmet_b200 import dmet as dmet_mod, fci_solver, localintegrals_hubbard as lh
from dmet_harness import CASES, CASES_2D, OracleDmet, build, chain_clusters


def test_fci_standin_known_answers():
    u = 4.0
    h = np.array([[0.0, -1.0], [-1.0, 0.0]])
    tei = np.zeros((2, 2, 2, 2))
    tei[0, 0, 0, 0] = tei[1, 1, 1, 1] = u
    e, d = fci_solver.solve(0.0, h, h, tei, 2, 2, 2)
    assert abs(e - (u / 2 - np.sqrt((u / 2) ** 2 + 4.0))) < 1e-12      # Hubbard dimer ground state
    assert abs(np.trace(d) - 2.0) < 1e-12
    # non-interacting 4-orbital system: energy = 2 * sum of the two lowest levels
    h4 = lh.hubbard_chain(4)
    e4, d4 = fci_solver.solve(0.0, h4, h4, np.zeros((4, 4, 4, 4)), 4, 4, 4)
    assert abs(e4 - 2 * np.sort(np.linalg.eigvalsh(h4))[:2].sum()) < 1e-12


def test_fci_davidson_route_matches_dense_route(monkeypatch):
    """The Kronecker-sum Davidson route (7..10 orbitals) against the dense route on a 6-orbital problem with a
    dense two-electron tensor, then an 8-orbital known answer (non-interacting: sum of the lowest levels)."""
    rng = np.random.default_rng(5)
    n = 6
    h = rng.standard_normal((n, n))
    h = h + h.T
    x = 0.3 * rng.standard_normal((n, n, 3))
    x = x + x.transpose(1, 0, 2)
    tei = np.einsum('ijx,klx->ijkl', x, x)                               # (ij|kl) with the 8-fold symmetry
    e_dense, d_dense = fci_solver.solve(0.1, h, h, tei, n, 6, 3, chempot_imp=0.2)
    monkeypatch.setattr(fci_solver, "DENSE_MAX_ORBITALS", 2)
    e_dav, d_dav = fci_solver.solve(0.1, h, h, tei, n, 6, 3, chempot_imp=0.2)
    assert abs(e_dav - e_dense) < 1e-10
    assert np.abs(d_dav - d_dense).max() < 1e-8
    monkeypatch.undo()
    h8 = lh.hubbard_chain(8) + np.diag(0.05 * np.arange(8))
    e8, d8 = fci_solver.solve(0.0, h8, h8, np.zeros((8,) * 4), 8, 8, 8)
    assert abs(e8 - 2 * np.sort(np.linalg.eigvalsh(h8))[:4].sum()) < 1e-9
    assert abs(np.trace(d8) - 8.0) < 1e-9


def test_square_lattice_loop_runs_with_8_orbital_impurity_problem():
    """configs[1] shape on CPU: 2 x 2 impurity + 4 bath orbitals (FCI dimension 4900) on the 8 x 8 lattice."""
    d = build("square8_TI_imp_only", OracleDmet)
    e = d.doselfconsistent()
    assert d.dmetOrbs[0].shape == (64, 8)
    assert e < d.ints.fullEhf - 1.0
    assert np.abs(d.umat[:4, :4]).max() > 0.1 and abs(np.trace(d.umat)) < 1e-9


def test_bath_construction_properties():
    ints = lh.localintegrals_hubbard(lh.hubbard_chain(12), 4.0, 12)
    d = OracleDmet(ints, chain_clusters(12, 2), False)
    one = d.helper.construct1RDM_loc(False, d.umat)
    nb, loc2dmet, core = d.helper.constructbath(one, d.impClust[1], 2)
    assert nb == 2
    assert np.allclose(loc2dmet.T @ loc2dmet, np.eye(12), atol=1e-12)
    assert np.allclose(loc2dmet[[2, 3], :2], np.eye(2))                 # impurity orbitals first
    rot = loc2dmet.T @ one @ loc2dmet
    assert np.abs(rot[4:, :4]).max() < 1e-10                             # environment decouples (Schmidt)
    assert np.all((np.abs(core[4:]) < 1e-10) | (np.abs(core[4:] - 2) < 1e-10))
    assert abs(np.trace(rot[:4, :4]) - 4.0) < 1e-10                      # half-filled embedded problem


def test_noninteracting_limit_is_exact():
    """U = 0: the impurity 1-RDM equals the mean-field one, so umat stays zero and E = E_RHF."""
    ints = lh.localintegrals_hubbard(lh.hubbard_chain(12), 0.0, 12)
    d = OracleDmet(ints, chain_clusters(12, 2), False, fitImpBath=False)
    e = d.doselfconsistent()
    assert abs(e - ints.fullEhf) < 1e-9
    assert np.abs(d.umat).max() < 1e-7
    assert d.iterations == 1


def test_oracle_loop_converges_and_lowers_energy():
    d = build("chain12_U4_imp_only", OracleDmet)
    e = d.doselfconsistent()
    assert d.iterations < 50
    assert e < d.ints.fullEhf - 1.0                                       # correlation energy recovered
    assert abs(np.trace(d.umat)) < 1e-9                                   # chemical-potential gauge removed
    assert abs(d.numeleccostfunction(d.mu_imp)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_converged_energy_matches_oracle_loop(case):
    ref = build(case, OracleDmet)
    e_ref = ref.doselfconsistent()
    gpu = build(case, dmet_mod.dmet)
    e_gpu = gpu.doselfconsistent()
    assert gpu.iterations == ref.iterations
    assert abs(e_gpu - e_ref) <= 1e-8                                     # 1e-8 Hartree
    assert np.abs(gpu.umat - ref.umat).max() <= 1e-7
    assert gpu.time_grad > 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES_2D))
def test_converged_energy_matches_oracle_loop_square_lattice(case):
    """2-D Hubbard with 2 x 2 impurities (BASELINE.json configs[1]: 24 x 24 included): the 8-orbital embedded
    problems go through the Davidson route of the FCI stand-in on both sides."""
    ref = build(case, OracleDmet)
    e_ref = ref.doselfconsistent()
    gpu = build(case, dmet_mod.dmet)
    e_gpu = gpu.doselfconsistent()
    assert ref.iterations < 50 and gpu.iterations == ref.iterations
    assert abs(e_gpu - e_ref) <= 1e-8                                     # 1e-8 Hartree
    assert np.abs(gpu.umat - ref.umat).max() <= 1e-7
    assert gpu.dmetOrbs[0].shape[1] == 8


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["chain12_U4_imp_only", "chain8_U2_bath", "chain12_U4_TI_bath"])
def test_gpu_levenberg_marquardt_reaches_the_same_fit(case):
    """f-3: for one DMET iteration (fixed impurity 1-RDMs) the device-side LM (Jacobian never leaves HBM)
    finds the same minimiser as scipy.optimize.leastsq (MINPACK) run to tight tolerances, up to the
    chemical-potential gauge that src/dmet.py:636 removes."""
    from scipy import optimize
    d = build(case, dmet_mod.dmet)
    d.mu_imp = optimize.newton(d.numeleccostfunction, 0.0)
    x0 = d.square2flat(d.umat)
    ref = optimize.leastsq(d.rdm_differences, x0, Dfun=d.rdm_differences_derivative, factor=0.1,
                           ftol=1e-15, xtol=1e-15, gtol=0.0)[0]
    got = d.leastsq_gpu(x0)
    gauge = lambda u: u - np.eye(u.shape[0]) * np.average(np.diag(u))
    u_ref, u_got = gauge(d.flat2square(ref)), gauge(d.flat2square(got))
    assert abs(d.costfunction(got) - d.costfunction(ref)) <= 1e-12
    assert np.abs(u_got - u_ref).max() <= 1e-6
    assert d.lm_iterations >= 1
    # whole loop: both optimisers stop the outer iteration at |du| < 1e-5 (src/dmet.py:600), so the
    # energies agree to that level
    a = build(case, dmet_mod.dmet)
    e_ref = a.doselfconsistent()
    b = build(case, dmet_mod.dmet)
    b.fit_on_gpu = True
    e_gpu = b.doselfconsistent()
    assert abs(e_gpu - e_ref) <= 5e-5


@pytest.mark.gpu
def test_lm_step_matches_numpy_normal_equations():
    from qcdmet_b200 import _lib, workloads as wl
    from oracle import response_oracle as ro
    w = wl.sweep(64)
    e, c = ro.eigh_sorted(w.h0)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_orbitals(c, e)
    plan.compute()
    jac = plan.jacobian()
    rng = np.random.default_rng(2)
    r = rng.standard_normal(w.nresid)
    for lam in (1e-2, 1.0):
        step, gmax, gdots = plan.lm_step(r, lam)
        a = jac.T @ jac
        g = jac.T @ r
        ref = -np.linalg.solve(a + lam * np.diag(np.maximum(np.diag(a), 1e-12)), g)
        assert np.abs(step - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
        assert abs(gmax - np.abs(g).max()) <= 1e-10 * np.abs(g).max()
        assert abs(gdots - g @ ref) <= 1e-8 * abs(g @ ref)
    plan.set_h0(w.h0)                       # new orbitals: the Jacobian in memory still serves LM retries
    step2, *_ = plan.lm_step(r, 1.0)
    assert np.array_equal(step2, step)
    plan.close()


@pytest.mark.gpu
def test_lm_step_qr_route_matches_lstsq_and_survives_ill_conditioning():
    """QCDMET_LM_QR factorises [J; sqrt(lambda) D] (cuSOLVER geqrf), as MINPACK's lmpar does, instead of forming J^T J.
    Same step as the normal-equation route on a well-conditioned fit; on a Jacobian whose columns are scaled over 12
    decades (cond(J)^2 far beyond double precision) the step still matches a QR-based host solve."""
    from qcdmet_b200 import _lib, workloads as wl
    from oracle import response_oracle as ro
    w = wl.sweep(64)
    e, c = ro.eigh_sorted(w.h0)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_orbitals(c, e)
    plan.compute()
    jac = plan.jacobian()
    rng = np.random.default_rng(4)
    r = rng.standard_normal(w.nresid)

    def host_step(j, lam):
        d = np.sqrt(np.maximum(np.sum(j * j, axis=0), 1e-12))
        aug = np.vstack([j, np.sqrt(lam) * np.diag(d)])
        rhs = np.concatenate([-r, np.zeros(j.shape[1])])
        return np.linalg.lstsq(aug, rhs, rcond=None)[0]

    for lam in (1e-3, 1.0):
        plan.set_lm_method(_lib.LM_NORMAL)
        s_normal, gmax_n, _ = plan.lm_step(r, lam)
        plan.set_lm_method(_lib.LM_QR)
        s_qr, gmax_q, gdots = plan.lm_step(r, lam)
        ref = host_step(jac, lam)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(s_qr - ref).max() <= 1e-9 * scale
        assert np.abs(s_qr - s_normal).max() <= 1e-7 * scale
        assert abs(gmax_q - gmax_n) <= 1e-12 * gmax_n
        assert abs(gdots - (jac.T @ r) @ ref) <= 1e-7 * abs((jac.T @ r) @ ref)
    plan.close()


@pytest.mark.gpu
def test_gpu_lm_reaches_the_minpack_fit_on_a_square_lattice():
    """f-3 on the configs[1] kind of fit (2-D Hubbard, 2 x 2 impurities, non-TI, impurity-only rows; 16 x 16 lattice:
    P = 640 parameters, Q = 1024 residuals): device Levenberg-Marquardt, both routes, against MINPACK."""
    from scipy import optimize
    from qcdmet_b200 import _lib, workloads as wl
    l = 16
    n = l * l
    # staggered on-site potential: a gapped mean field, so that no level crosses the Fermi energy along the fit (on the
    # bare lattice the residual is discontinuous there and every optimiser's path depends on where it first steps)
    stagger = np.array([0.4 if (k % 4) in (0, 3) else -0.4 for k in range(n)])
    ints = lh.localintegrals_hubbard(wl.hubbard2d_hopping(l, l) + np.diag(stagger), 4.0, n)
    d = dmet_mod.dmet(ints, chain_clusters(n, 4), False, method='ED', SCmethod='LSTSQ', fitImpBath=False)
    # synthetic impurity 1-RDMs: the mean-field blocks with the nearest-neighbour coherences reduced (what the
    # correlated solver does), so that the fit has a non-trivial, well-defined minimiser
    mf = ints.origDMloc
    d.imp_1RDM = []
    for f in range(n // 4):
        blk = mf[4 * f:4 * f + 4, 4 * f:4 * f + 4].copy()
        blk[~np.eye(4, dtype=bool)] *= 0.8
        d.imp_1RDM.append(blk)
    x0 = np.zeros(d.helper.Nterms)
    ref = optimize.leastsq(d.rdm_differences, x0, Dfun=d.rdm_differences_derivative, factor=0.1,
                           ftol=1e-15, xtol=1e-15, gtol=0.0)[0]
    gauge = lambda u: u - np.eye(n) * np.average(np.diag(u))
    for method in (_lib.LM_NORMAL, _lib.LM_QR):
        d.lm_method = method
        got = d.leastsq_gpu(x0)
        assert abs(d.costfunction(got) - d.costfunction(ref)) <= 1e-11
        assert np.abs(gauge(d.flat2square(got)) - gauge(d.flat2square(ref))).max() <= 1e-6
        assert d.lm_iterations >= 2
