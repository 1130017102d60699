import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import response_oracle as ro
from qcdmet_b200 import _lib, workloads as wl
print("lib", _lib.LIB_PATH, "resident env", os.environ.get("QCDMET_B200_CONTRACT_RESIDENT"), "slots", os.environ.get("QCDMET_B200_CONTRACT_SLOTS"))
for name, w in (("frag14", wl.make_workload("frag14", 56, 12, [14] * 4)), ("sn2s", wl.make_workload("sn2s", 150, 30, [52, 24, 29, 24, 21], nterms=300)),
                ("polyyne", wl.polyyne_like())):
    e, c = ro.eigh_sorted(w.h0)
    ref = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    for rep in range(3):
        plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
        plan.set_kernel(_lib.KERNEL_FACTORISED)
        plan.set_orbitals(c, e)
        plan.compute()
        jac = plan.jacobian()
        err = np.abs(jac - ref)
        bad = np.argwhere(err > 1e-10)
        print(name, rep, "max err", err.max(), "bad entries", len(bad), "first", bad[:4].tolist(), "nan", int(np.isnan(jac).sum()))
        plan.close()
