"""GPU parity at the sizes BASELINE.json's configs STATE, against the unmodified reference kernel compiled
into oracle/_ref (its own LAPACK eigensolve and response loop; Jacobian assembled by src/dmet.py:496-520):
configs[1] 24x24 Hubbard with 2x2 impurities at l = 24 (every column), the sn2 shape on 500+ spread columns,
and the north-star sweep (Norb 4000, P 10^4) on 32 spread columns.  Tolerance 1e-10 absolute."""
import numpy as np
import pytest

from qcdmet_b200 import _lib, workloads as wl
from util import ref_available, ref_jacobian_columns

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")]
TOL = 1e-10


def gpu_jacobian(w, kernel=_lib.KERNEL_AUTO):
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_kernel(kernel)
    plan.set_h0(w.h0)                   # cuSOLVER syevd on the device: the whole path, not given orbitals
    plan.compute()
    jac, used = plan.jacobian(), plan.kernel_used
    plan.close()
    return jac, used


def spread(p, n):
    return np.unique(np.linspace(0, p - 1, n).round().astype(np.int64))


def test_hubbard24_l24_every_column_against_compiled_reference():
    """configs[1]: Norb = 576, P = 1440, Q = 2304, non-TI, impurity-only rows -- all 1440 columns."""
    w = wl.hubbard24()
    assert (w.norb, w.nterms, w.nresid) == (576, 1440, 2304)
    jac, used = gpu_jacobian(w)
    assert used == _lib.KERNEL_DMMA
    ref = ref_jacobian_columns(w, np.arange(w.nterms), chunk=240)
    assert np.abs(jac - ref).max() <= TOL
    # the independent general kernel on the same full matrix
    gen, used = gpu_jacobian(w, _lib.KERNEL_GENERAL)
    assert used == _lib.KERNEL_GENERAL and np.abs(gen - ref).max() <= TOL


def test_hubbard24_l24_translation_invariant():
    """configs[1], TI list: P = 10 perturbations of 144 / 288 elements each (one per cell)."""
    for bath in (False, True):
        w = wl.hubbard24(trans_inv=True, fit_imp_bath=bath)
        assert w.nterms == 10 and int(np.diff(w.h1start).max()) == 288
        jac, used = gpu_jacobian(w)
        assert used == _lib.KERNEL_DMMA
        ref = ref_jacobian_columns(w, np.arange(10))
        assert np.abs(jac - ref).max() <= TOL


def test_hubbard24_l24_impurity_plus_bath_rows():
    """configs[1] with fitImpBath rows (Q = 144 x 64 = 9216): 160 columns spread over all 1440."""
    w = wl.hubbard24(fit_imp_bath=True)
    assert (w.nterms, w.nresid) == (1440, 9216)
    jac, used = gpu_jacobian(w)
    assert used == _lib.KERNEL_DMMA
    cols = spread(w.nterms, 160)
    ref = ref_jacobian_columns(w, cols, chunk=80)
    assert np.abs(jac[:, cols] - ref).max() <= TOL


def test_sn2_shape_spread_columns():
    """configs[3] shape (Norb 321, fragments [52, 24 x 10, 29], P = 4813): 512 columns spread over [0, P)."""
    w = wl.sn2_like()
    jac, used = gpu_jacobian(w)
    assert used == _lib.KERNEL_FACTORISED
    cols = spread(w.nterms, 512)
    assert cols[0] == 0 and cols[-1] == w.nterms - 1
    ref = ref_jacobian_columns(w, cols, chunk=256)
    assert np.abs(jac[:, cols] - ref).max() <= TOL


def test_polyyne_shape_every_column():
    """configs[2] shape (Norb 168, 12 fragments of 14, P = 1260): every column."""
    w = wl.polyyne_like()
    jac, used = gpu_jacobian(w)
    assert used == _lib.KERNEL_FACTORISED
    ref = ref_jacobian_columns(w, np.arange(w.nterms), chunk=420)
    assert np.abs(jac - ref).max() <= TOL


def test_north_star_spread_columns_against_compiled_reference():
    """configs[4] at the target: Norb = 4000, P = 10^4, Q = 16000.  32 columns spread over [0, P) from the
    compiled reference (each a full 4000 x 4000 dD_p from its own dsyev orbitals)."""
    w = wl.sweep(4000)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_h0(w.h0)
    plan.compute()
    assert plan.kernel_used == _lib.KERNEL_DMMA and plan.symmetric_used
    cols = spread(w.nterms, 32)
    got = np.stack([plan.jacobian_columns(int(c), int(c) + 1)[:, 0] for c in cols], axis=1)
    plan.close()
    ref = ref_jacobian_columns(w, cols, chunk=32)
    assert np.abs(got - ref).max() <= TOL
