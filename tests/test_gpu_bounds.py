"""Memory safety without compute-sanitizer (closed on the GPU pool): the bounds-checked build of the library
(make debug -> libqcdmet_b200_dbg.so, QC_DEV_ASSERT on every Jacobian / panel store and on the table-driven indices)
runs ragged and degenerate shapes of every kernel family in a fresh process; a violated check traps the kernel and
the process fails."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "qcdmet_b200", "libqcdmet_b200_dbg.so")

SCRIPT = r"""
import numpy as np
from oracle import response_oracle as ro
from qcdmet_b200 import _lib, workloads as wl
assert _lib.LIB_PATH.endswith("_dbg.so"), _lib.LIB_PATH
cases = [
    (wl.sweep(128), _lib.KERNEL_AUTO),
    (wl.make_workload("ragged", 70, 31, [4, 3, 4, 2, 4, 4, 1, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4]), _lib.KERNEL_AUTO),
    (wl.hubbard24(fit_imp_bath=True, l=8), _lib.KERNEL_AUTO),
    (wl.hubbard24(trans_inv=True, l=8), _lib.KERNEL_AUTO),
    (wl.make_workload("frag14", 56, 12, [14] * 4), _lib.KERNEL_AUTO),
    (wl.make_workload("sn2s", 150, 30, [52, 24, 29, 24, 21], nterms=300), _lib.KERNEL_FACTORISED),
    (wl.make_workload("big", 150, 70, [100, 30, 9], nterms=200), _lib.KERNEL_FACTORISED),
    (wl.sweep(64), _lib.KERNEL_GENERAL),
]
for w, kernel in cases:
    e, c = ro.eigh_sorted(w.h0)
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    plan.set_kernel(kernel)
    plan.set_orbitals(c, e)
    plan.compute()
    jac = plan.jacobian()
    ref = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    assert np.abs(jac - ref).max() <= 1e-10, w.name
    if w.rows is None and int(np.diff(w.row_start).max()) <= 64:
        nb, occ, bath = plan.construct_baths(w.row_start, w.row_idx)
        assert np.isfinite(bath).all()
    plan.set_columns(3, min(w.nterms, 41))
    plan.compute()
    assert np.abs(plan.jacobian() - ref[:, 3:min(w.nterms, 41)]).max() <= 1e-10, w.name
    plan.close()
print("bounds-checked build: ok")
"""


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(DBG), reason="bounds-checked library not built (make -C qcdmet_b200/csrc debug)")
def test_bounds_checked_build_runs_clean():
    env = dict(os.environ, QCDMET_B200_LIB=DBG, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "bounds check failed" not in out.stdout + out.stderr
    assert "bounds-checked build: ok" in out.stdout
