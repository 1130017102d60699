"""Smaller pieces of the reference API on the fit path: the natural-orbital rotation plumbing
(src/dmet.py:343-359, 482, 585-586), verify_gradient / hessian_eigenvalues (:527-558, :622-623) and the
report a plan gives when AUTO has to leave the tensor-core kernels."""
import numpy as np
import pytest

from oracle import response_oracle as ro
from qcdmet_b200 import _lib, dmet as dmet_mod, localintegrals_hubbard as lh, workloads as wl
from dmet_harness import OracleDmet, build, chain_clusters

pytestmark = pytest.mark.gpu


def random_orthogonal(m, rng):
    q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    return q


@pytest.mark.parametrize("trans_inv", [False, True])
def test_norotation_jacobian_matches_restated_reference(trans_inv):
    """With a natural-orbital rotation R the reference diagonalises R^T (Fock + R u R^T) R and takes the impurity
    blocks of the rotated-basis response (src/qcdmethelper.py:98-99, src/dmet.py:482,509,585-586)."""
    n, m = 16, 2
    ints = lh.localintegrals_hubbard(lh.hubbard_chain(n), 4.0, n)
    d = dmet_mod.dmet(ints, chain_clusters(n, m), trans_inv, method='ED', SCmethod='LSTSQ', fitImpBath=False)
    rng = np.random.default_rng(8)
    d.NOvecs = [random_orthogonal(m, rng) for _ in range(1 if trans_inv else n // m)]
    d.NOrotation = d.constructNOrotation()
    assert np.abs(d.NOrotation.T @ d.NOrotation - np.eye(n)).max() < 1e-12
    if trans_inv:
        assert np.array_equal(d.NOrotation[2:4, 2:4], d.NOrotation[:2, :2])
    x = 0.05 * rng.standard_normal(d.helper.Nterms)
    u = d.flat2square(x)
    assert np.abs(d.square2flat(u) - x).max() < 1e-12        # the two maps are inverse to each other
    jac = d.rdm_differences_derivative(x)
    hp = d.helper
    resp = ro.construct1rdm_response(ints.loc_rhf_fock(), u, hp.numPairs, hp.H1start, hp.H1row, hp.H1col,
                                     no_rotation=d.NOrotation)
    cols = []
    for p in range(hp.Nterms):
        blocks = []
        for f in range(len(d.imp_size)):
            idx = np.flatnonzero(d.impClust[f])
            blocks.append(resp[p][np.ix_(idx, idx)].reshape(-1, order="F"))
        cols.append(np.concatenate(blocks))
    ref = np.array(cols).T
    assert jac.shape == ref.shape
    assert np.abs(jac - ref).max() <= 1e-10


def test_verify_gradient_and_hessian_spectrum():
    """verify_gradient / hessian_eigenvalues on the device path against the same formulas driven by the CPU
    oracle, and the flag that runs the Hessian inside doselfconsistent as the reference does (:622-623)."""
    gpu = build("chain8_U2_bath", dmet_mod.dmet)
    cpu = build("chain8_U2_bath", OracleDmet)
    for d in (gpu, cpu):
        d.mu_imp = 0.0
        d.doexact(0.0)                                        # impurity 1-RDMs and dmetOrbs of the first iteration
    rng = np.random.default_rng(3)
    x = 0.02 * rng.standard_normal(gpu.helper.Nterms)
    diff, norm = gpu.verify_gradient(x)
    assert norm > 1e-3 and diff <= 1e-5 * max(1.0, norm)      # forward differences with step 1e-7
    h_gpu = gpu.hessian_eigenvalues(x)
    g0 = 2.0 * cpu.rdm_differences_derivative(x).T @ cpu.rdm_differences(x)
    cols = []
    for k in range(x.size):
        xk = x.copy()
        xk[k] += 1e-7
        cols.append((2.0 * cpu.rdm_differences_derivative(xk).T @ cpu.rdm_differences(xk) - g0) / 1e-7)
    hess = np.stack(cols, axis=1)
    h_cpu = np.linalg.eigvalsh(0.5 * (hess + hess.T))
    assert np.abs(h_gpu - h_cpu).max() <= 1e-4 * max(1.0, np.abs(h_cpu).max())
    loop = build("chain8_U2_bath", dmet_mod.dmet)
    loop.print_hessian = True
    loop.doselfconsistent(max_iterations=2)
    assert loop.last_hessian_eigenvalues is not None and loop.last_hessian_eigenvalues.size == loop.helper.Nterms


def test_auto_reports_why_it_left_the_tensor_core_kernels():
    """Translation-invariant list on 40 cells of 5 orbitals: 80 elements per perturbation exceed what the fused and
    the factorised path stage; AUTO runs the general kernel and says so through qcdmet_plan_info."""
    w = wl.make_workload("ti5_long", 200, 90, [5] * 40, trans_inv=True)
    assert int(np.diff(w.h1start).max()) > 64
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_h0(w.h0)
    plan.compute()
    assert plan.kernel_used == _lib.KERNEL_GENERAL
    msg = plan.info
    assert "general" in msg and "64 elements" in msg
    c, e = plan.orbitals(), plan.eigenvalues()
    ref = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    assert np.abs(plan.jacobian() - ref).max() <= 1e-10
    plan.close()
    ok = wl.sweep(64)
    plan = _lib.Plan(ok.norb, ok.num_pairs, ok.h1start, ok.h1row, ok.h1col, ok.row_start, ok.rows, ok.row_idx)
    plan.set_h0(ok.h0)
    plan.compute()
    assert plan.info == ""                                  # nothing to report on the fused path
    plan.close()
