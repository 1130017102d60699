"""Single-process multi-GPU plans (qcdmet_plan_set_devices / qcdmet_set_devices): the in-library parallel
region that replaces the OpenMP team of lib/rhf_response.cc:73-79 behind the unchanged ctypes boundary.

On a one-GPU box the group members wrap around the visible devices (QCDMET_B200_OVERSUBSCRIBE=1): the same
code path -- one eigensolve + peer copies of (C, e), per-member streams, event barrier, fused peer stores or
column shards, threaded device->host copies -- with several members per device.  With >= 2 GPUs the same
tests also run on distinct devices."""
import os

import numpy as np
import pytest

from oracle import response_oracle as ro
from qcdmet_b200 import _lib, dmet as dmet_mod, workloads as wl

pytestmark = pytest.mark.gpu
TOL = 1e-10


def ndev():
    return _lib.load().qcdmet_device_count()


@pytest.fixture
def oversubscribe(monkeypatch):
    monkeypatch.setenv("QCDMET_B200_OVERSUBSCRIBE", "1")


def single_and_group(w, n, kernel=_lib.KERNEL_AUTO, cols=None):
    args = (w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    one = _lib.Plan(*args)
    one.set_kernel(kernel)
    if cols:
        one.set_columns(*cols)
    one.set_h0(w.h0)
    one.compute()
    ref = one.jacobian()
    c, e, d = one.orbitals(), one.eigenvalues(), one.dblocks()
    used = one.kernel_used
    one.close()
    grp = _lib.Plan(*args, devices=n)
    assert grp.num_devices == n
    grp.set_kernel(kernel)
    if cols:
        grp.set_columns(*cols)
    grp.set_h0(w.h0)
    grp.compute()
    got = grp.jacobian()
    assert grp.kernel_used == used
    assert np.array_equal(grp.orbitals(), c) and np.array_equal(grp.eigenvalues(), e) and np.array_equal(grp.dblocks(), d)
    return ref, got, grp


CASES = {
    "sweep_fused": (lambda: wl.sweep(128), _lib.KERNEL_AUTO),
    "sweep_ragged_fused": (lambda: wl.make_workload("r", 70, 31, [4, 3, 4, 2, 4, 4, 1, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4]), _lib.KERNEL_AUTO),
    "bath_rows_fused": (lambda: wl.hubbard24(fit_imp_bath=True, l=8), _lib.KERNEL_AUTO),
    "ti_split_shards": (lambda: wl.hubbard24(trans_inv=True, l=8), _lib.KERNEL_AUTO),
    "frag14_factorised_shards": (lambda: wl.make_workload("frag14", 56, 12, [14] * 4), _lib.KERNEL_AUTO),
    "general_shards": (lambda: wl.sweep(64), _lib.KERNEL_GENERAL),
    "nosym_fused": (lambda: wl.sweep(96), _lib.KERNEL_DMMA_NOSYM),
}


@pytest.mark.parametrize("n", [2, 3])
@pytest.mark.parametrize("case", list(CASES))
def test_group_matches_single_device_bit_for_bit(case, n, oversubscribe):
    make, kernel = CASES[case]
    w = make()
    ref, got, grp = single_and_group(w, n, kernel)
    # same kernels, same per-element arithmetic, one owner per element: identical bits
    assert np.array_equal(got, ref)
    s, ss, bits, same = grp.checksum()
    assert same
    assert bits == int(np.bitwise_xor.reduce(got.view(np.uint64).reshape(-1)))
    assert abs(s - got.sum()) <= 1e-9 * max(1.0, np.abs(got).sum())
    # a second step on new orbitals through the same group (stream / event ordering between steps)
    h1 = w.h0 + 0.01 * np.diag(np.linspace(-1, 1, w.norb))
    grp.set_h0(h1)
    grp.compute()
    got2 = grp.jacobian(pinned=True)
    e, c = grp.eigenvalues(), grp.orbitals()
    ref2 = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    assert np.abs(got2 - ref2).max() <= TOL
    # column slices of the result, shipped by whichever member holds them
    lo, hi = 1, max(2, w.nterms - 2)
    assert np.array_equal(grp.jacobian_columns(lo, hi), got2[:, lo:hi])
    grp.close()


def test_group_partial_column_range_and_lm_step(oversubscribe):
    w = wl.sweep(96)
    ref, got, grp = single_and_group(w, 2, cols=(10, 150))
    assert got.shape == (w.nresid, 140) and np.array_equal(got, ref)
    grp.set_columns(0, w.nterms)
    grp.compute()
    jac = grp.jacobian()
    r = np.random.default_rng(3).standard_normal(w.nresid)
    step, gmax, gdots = grp.lm_step(r, 1e-2)          # J^T J from block rows computed by both members
    a, g = jac.T @ jac, jac.T @ r
    want = -np.linalg.solve(a + 1e-2 * np.diag(np.maximum(np.diag(a), 1e-12)), g)
    assert np.abs(step - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    assert abs(gmax - np.abs(g).max()) <= 1e-10 * np.abs(g).max()
    assert np.abs(grp.gradient(r) - 2 * g).max() <= 1e-10 * max(1.0, np.abs(g).max())
    grp.close()


def test_group_errors_and_state(oversubscribe):
    w = wl.sweep(32)
    args = (w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    grp = _lib.Plan(*args, devices=2)
    with pytest.raises(_lib.QcdmetError):
        grp.export_jacobian()                          # IPC export is the multi-process route
    with pytest.raises(_lib.QcdmetError):
        grp.compute()                                  # no orbitals yet
    bad = np.zeros((32, 32))                           # zero gap: refused, and nothing stale may be used afterwards
    with pytest.raises(_lib.QcdmetError) as ei:
        grp.set_h0(bad)
    assert ei.value.code == -6
    with pytest.raises(_lib.QcdmetError) as ei:
        grp.compute()
    assert ei.value.code == -5
    grp.set_h0(w.h0)
    grp.compute()
    one = _lib.Plan(*args)
    one.set_h0(w.h0)
    one.compute()
    assert np.array_equal(grp.jacobian(), one.jacobian())
    grp.set_devices(1)                                 # back to one device: still the same answer
    assert grp.num_devices == 1
    grp.compute()
    assert np.array_equal(grp.jacobian(), one.jacobian())
    one.close()
    grp.close()
    if ndev() == 1:
        with pytest.raises(_lib.QcdmetError):          # without the test hook more members than devices is an error
            os.environ.pop("QCDMET_B200_OVERSUBSCRIBE")
            _lib.Plan(*args, devices=2)


def test_failed_set_h0_leaves_no_stale_orbitals():
    """a rejected H0 must not leave panels of other orbitals behind (single-device plan)."""
    w = wl.sweep(32)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_h0(w.h0)
    plan.compute()
    good = plan.jacobian()
    with pytest.raises(_lib.QcdmetError):
        plan.set_h0(np.zeros((32, 32)))
    for call in (plan.compute, plan.dblocks, plan.eigenvalues):
        with pytest.raises(_lib.QcdmetError) as ei:
            call()
        assert ei.value.code == -5
    plan.set_h0(w.h0)
    plan.compute()
    assert np.array_equal(plan.jacobian(), good)
    plan.close()


def test_process_default_devices_drive_the_dropin_api(oversubscribe):
    """qcdmet_set_devices(n): the stateless C entry point and the Python mirror of the reference's classes use
    n devices without any change on the caller's side."""
    w = wl.hubbard24(l=8)
    lib = _lib.load()
    jac1 = np.empty((w.nresid, w.nterms), order="F")
    args = (w.norb, w.nterms, w.num_pairs, _lib._ptr(w.h1start), _lib._ptr(w.h1row), _lib._ptr(w.h1col),
            _lib._ptr(np.ascontiguousarray(w.h0)), len(w.row_start) - 1, _lib._ptr(w.row_start), None,
            _lib._ptr(w.row_idx), 0, w.nterms)
    _lib.check(lib.qcdmet_jacobian(*args, _lib._ptr(jac1), None, None))
    _lib.set_devices(2)
    try:
        assert _lib.get_devices() == 2
        jac2 = np.empty_like(jac1)
        _lib.check(lib.qcdmet_jacobian(*args, _lib._ptr(jac2), None, None))
        assert np.array_equal(jac1, jac2)

        class Ints:
            Norbs, Nelec, ERIinMEM, TI_OK = w.norb, 2 * w.num_pairs, False, True
            loc_rhf_fock = staticmethod(lambda: w.h0)
            loc_oei = staticmethod(lambda: w.h0)

        d = dmet_mod.dmet(Ints(), w.imp_clust, False, fitImpBath=False)
        jac3 = d.rdm_differences_derivative(np.zeros(w.nterms))
        plan = next(iter(d.helper._plans.values()))
        assert plan.num_devices == 2
        assert np.array_equal(jac3, jac1)
    finally:
        _lib.set_devices(1)
        lib.qcdmet_shutdown()


@pytest.mark.skipif(ndev() < 2, reason="needs two GPUs")
def test_group_on_two_real_devices():
    w = wl.sweep(1000, nterms=2500)
    ref, got, grp = single_and_group(w, 2)
    assert np.array_equal(got, ref)
    assert grp.checksum()[3]
    t = grp.timings()
    assert t["syevd_ms"] and t["kernel_ms"] and t["d2h_ms"] is not None
    grp.close()
